// filter.cu -- K6: per-row predicate -> 1 bit/row allow-bitmap + exact candidate count.
//
// Replaces the WHERE clause the reference builds in app/retrieve.py:93-120
// (_build_filter_clause) plus `AND embedding IS NOT NULL` (app/retrieve.py:318,347), evaluated
// per row inside Postgres, and the exact COUNT(*) of app/retrieve.py:303-323
// (_estimate_dense_candidates).  HBM-bound column scan: 20 B/row read (call_slot i32,
// started_at i64, tag_bits u64) + 1 bit/row written; coalesced, one ballot per 32 rows.
#include "common.cuh"

namespace {

struct FilterParams {
    const int32_t *call_slot;
    const int64_t *started_at;
    const uint64_t *tag_bits;
    const uint32_t *valid;
    const uint32_t *call_bitmap;   // device copy, nullable
    int64_t n_call_slots;
    int64_t n_rows;
    int64_t n_words;               // words of the output bitmap = ceil(capacity / 32)
    int has_from, has_to, has_tags;
    int64_t date_from, date_to;
    uint64_t tag_any;
    uint32_t *out_allow;
    unsigned long long *out_count;
};

__global__ void __launch_bounds__(256) filter_bitmap_kernel(const FilterParams p)
{
    // The bitmap covers the store's CAPACITY (p.n_words words): rows appended after this launch read as "not
    // allowed" instead of lying beyond the bitmap of a scan that sees the larger row count.
    const int64_t words = p.n_words;
    unsigned long long local = 0;
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    // Four bitmap words (128 rows) per warp and trip, every column read issued before the first is used: the kernel is a
    // short chain of dependent loads per row (validity word -> call slot -> call bitmap), so its time is load latency x
    // trips, not bytes (12 us for 1 M rows with one word per trip against 3 us of column traffic).
    constexpr int U = 4;
    for (int64_t w0 = warp0 * U; w0 < words; w0 += nwarps * U) {
        uint32_t vbits[U];
        int32_t slot[U];
        int64_t ts[U];
        uint64_t tags[U];
        bool in[U];
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const int64_t w = w0 + j, row = (w << 5) + lane;
            in[j] = w < words && row < p.n_rows;
            vbits[j] = in[j] ? p.valid[w] : 0u;
            slot[j] = (in[j] && p.call_bitmap) ? p.call_slot[row] : 0;
            ts[j] = (in[j] && (p.has_from | p.has_to)) ? p.started_at[row] : 0;
            tags[j] = (in[j] && p.has_tags) ? p.tag_bits[row] : 0ull;
        }
        uint32_t cword[U];
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const bool s_ok = p.call_bitmap && slot[j] >= 0 && slot[j] < p.n_call_slots;
            cword[j] = s_ok ? p.call_bitmap[slot[j] >> 5] : 0u;
        }
#pragma unroll
        for (int j = 0; j < U; ++j) {
            bool ok = in[j] && ((vbits[j] >> lane) & 1u);
            if (p.call_bitmap) ok = ok && ((cword[j] >> (slot[j] & 31)) & 1u);
            if (p.has_from) ok = ok && ts[j] >= p.date_from;
            if (p.has_to) ok = ok && ts[j] <= p.date_to;
            if (p.has_tags) ok = ok && (tags[j] & p.tag_any) != 0ull;
            const unsigned bits = __ballot_sync(0xffffffffu, ok);
            if (lane == 0 && w0 + j < words) {
                p.out_allow[w0 + j] = bits;
                local += __popc(bits);
            }
        }
    }
    // one atomic per block
    __shared__ unsigned long long s_cnt[8];
    if (lane == 0) s_cnt[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += s_cnt[i];
        if (t) atomicAdd(p.out_count, t);
    }
}

}  // namespace

// Enqueue K6 on `st`.  bm_dev: device copy of the call-slot bitmap (nullable); cnt_dev must be zero.
int cdr_filter_launch(cdr_store *s, const uint32_t *bm_dev, int64_t n_call_slots, int has_from, int64_t date_from_us,
                      int has_to, int64_t date_to_us, int has_tags, uint64_t tag_any, uint32_t *out_allow_dev,
                      unsigned long long *cnt_dev, cudaStream_t st)
{
    FilterParams p;
    p.call_slot = s->call_slot;
    p.started_at = s->started_at;
    p.tag_bits = s->tag_bits;
    p.valid = s->valid;
    p.call_bitmap = bm_dev;
    p.n_call_slots = n_call_slots;
    p.n_rows = s->n_rows;
    p.n_words = (s->capacity + 31) / 32;
    p.has_from = has_from != 0;
    p.has_to = has_to != 0;
    p.has_tags = has_tags != 0;
    p.date_from = date_from_us;
    p.date_to = date_to_us;
    p.tag_any = tag_any;
    p.out_allow = out_allow_dev;
    p.out_count = cnt_dev;
    const int64_t words = p.n_words;
    int64_t blocks = (words + 7) / 8;
    const int64_t cap = (int64_t)s->sm_count * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    filter_bitmap_kernel<<<(unsigned)blocks, 256, 0, st>>>(p);
    CDR_LAUNCH_CHECK();
    return CDR_OK;
}

extern "C" int32_t cdr_filter_build(cdr_store *s, const uint32_t *call_slot_bitmap_host,
                                    int64_t n_call_slots, int32_t has_date_from, int64_t date_from_us,
                                    int32_t has_date_to, int64_t date_to_us, int32_t has_tag_filter,
                                    uint64_t tag_any, uint32_t *out_allow_dev, int64_t *out_count_host,
                                    void *stream)
{
    CDR_REQUIRE(s != nullptr && out_allow_dev != nullptr, CDR_ERR_INVALID, "cdr_filter_build: NULL argument");
    CDR_REQUIRE(s->finalized, CDR_ERR_STATE, "cdr_filter_build: store not finalized");
    CDR_REQUIRE(call_slot_bitmap_host == nullptr || n_call_slots >= 0, CDR_ERR_INVALID,
                "cdr_filter_build: n_call_slots < 0");
    DeviceGuard g(s->device);
    cudaStream_t st = (cudaStream_t)stream;
    // this thread's device staging (a stream-ordered pool allocation per call costs milliseconds once the pool
    // has trimmed at the previous synchronisation): [counter | call-slot bitmap]
    const size_t bm_words = call_slot_bitmap_host ? (size_t)((n_call_slots + 31) / 32) + 1 : 0;
    unsigned char *stage = (unsigned char *)cdr_thread_device(s->device, 16 + bm_words * 4);
    if (!stage) return CDR_ERR_OOM;
    unsigned long long *cnt = reinterpret_cast<unsigned long long *>(stage);
    uint32_t *bm_dev = call_slot_bitmap_host ? reinterpret_cast<uint32_t *>(stage + 16) : nullptr;
    CDR_CUDA(cudaMemsetAsync(stage, 0, 16 + bm_words * 4, st));
    if (bm_dev && n_call_slots > 0)
        CDR_CUDA(cudaMemcpyAsync(bm_dev, call_slot_bitmap_host, (size_t)((n_call_slots + 31) / 32) * 4,
                                 cudaMemcpyHostToDevice, st));
    int rc;
    {
        std::lock_guard<std::mutex> lk(s->mu);          // one row count for the whole bitmap (appends take the same lock)
        rc = cdr_filter_launch(s, bm_dev, n_call_slots, has_date_from, date_from_us, has_date_to, date_to_us,
                               has_tag_filter, tag_any, out_allow_dev, cnt, st);
    }
    if (rc != CDR_OK) return rc;
    unsigned long long h = 0;
    CDR_CUDA(cudaMemcpyAsync(&h, cnt, sizeof(h), cudaMemcpyDeviceToHost, st));
    CDR_CUDA(cudaStreamSynchronize(st));
    if (out_count_host) *out_count_host = (int64_t)h;
    return CDR_OK;
}
