// tech_lane.cu -- device-resident tech_tokens lexical lane.
//
// Replaces the SQL of _fetch_chunks_tech / _fetch_artifacts_tech (app/retrieve.py:183-242):
//     WHERE <filters> AND tech_tokens && :tokens
//     ORDER BY call_started_at DESC, <id> ASC LIMIT :limit
// which Postgres answers from a GIN index on the text[] column (alembic 0001:96).  Here the index
// is a dictionary-encoded CSR posting structure in HBM (token id -> ascending row list) plus one
// precomputed u32 per row: its rank in the global order (call_started_at DESC, id ASC).  The lane
// has no score, so "top-limit" = the `limit` smallest ranks among matching rows: the same packed
// u64 key / warp top-k machinery as the dense lane applies, with key =
// ((0xFFFFFFFF - rank) << 32) | (0xFFFFFFFF - row).
//
// Posting lists are kept in RANK order (best row of the lane's ORDER BY first; re-sorted at index build), with the
// rank stored beside the row.  One CTA (8 warps) per query; warp w takes the 32-posting chunks w, w+8, ... of every
// query token.  Candidates therefore arrive best-first and in order: the postings that pass the filter are compacted
// into an already-sorted staged list (no insertion per posting) and the warp leaves the token once KC are staged or
// the chunk's best possible key cannot enter its running list (unfiltered: after 2 chunks per warp, whatever the list
// length).  Under a selective filter the warps keep scanning until enough rows pass.  The filter predicate
// (call_id = ANY, date range, tags overlap -- app/retrieve.py:93-120, WITHOUT the dense lane's `embedding IS NOT
// NULL`) is evaluated on the posting rows only.  A row matching several tokens appears in several lists: every merge
// (staged -> running list, and the pairwise tree over the 8 warps) drops identical keys before it cuts to KC.
// (Round 1/2 history: a scan in row order paid an insertion per posting, 4.6 ms for 64 queries over frequent tokens;
// rank order with one-by-one insertion + a 512-key shared-memory sort cost 98 us per single-query launch.)
#include "common.cuh"

#include <cooperative_groups.h>

#include <algorithm>
#include <cstdlib>
#include <vector>

struct cdr_tech_index {
    cdr_store *store = nullptr;
    int32_t n_tokens = 0;
    int64_t n_postings = 0;
    int64_t *offsets = nullptr;    // [n_tokens + 1]
    uint32_t *rows = nullptr;      // [n_postings]  rows of a token in rank order
    uint32_t *prank = nullptr;     // [n_postings]  rank of each posting's row (ascending inside a token)
};

namespace {

constexpr int kTechWarps = 8;
constexpr int kTechMaxTokens = 32;

struct TechParams {
    const int64_t *offsets;
    const uint32_t *post_rows;
    const uint32_t *post_rank;
    const int64_t *ids;
    const int32_t *call_slot;
    const int64_t *started_at;
    const uint64_t *tag_bits;
    const uint32_t *call_bitmap;   // nullable
    int64_t n_call_slots;
    int64_t n_rows;
    int n_index_tokens;
    int has_from, has_to, has_tags;
    int64_t date_from, date_to;
    uint64_t tag_any;
    const int32_t *q_tokens;       // [nq, max_tokens]
    const int32_t *q_ntok;         // [nq]
    int max_tokens;
    int limit;
    int64_t *out_ids;              // [nq, limit]
    int32_t *out_n;                // [nq]
};

// a: this warp's running list (registers, sorted descending, distinct, CDR_EMPTY_KEY = 0 at the tail); b: another such
// list of KC keys in shared memory.  Result in a: the first KC DISTINCT keys of the union, sorted.  Identical keys (a row
// that matched several tokens) are dropped BEFORE the cut to KC, so duplicates never push a row out.  scratch: KC keys.
template <int NPL>
__device__ __forceinline__ void warp_merge_distinct(uint64_t (&a)[NPL], const uint64_t *b, uint64_t *scratch, int lane)
{
    constexpr int KC = NPL * 32;
    uint64_t m[2 * NPL];
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
        m[i] = a[i];
        m[NPL + i] = b[KC - 1 - (i * 32 + lane)];          // ascending half: the 2 KC keys form a bitonic sequence
    }
    warp_bitonic_merge_desc<2 * NPL>(m, lane);
    int off = 0;
#pragma unroll
    for (int i = 0; i < 2 * NPL; ++i) {
        const uint64_t up = __shfl_up_sync(0xffffffffu, m[i], 1);
        const uint64_t wrap = i > 0 ? __shfl_sync(0xffffffffu, m[i > 0 ? i - 1 : 0], 31) : ~0ull;
        const uint64_t prev = lane > 0 ? up : wrap;
        const bool keep = m[i] != CDR_EMPTY_KEY && m[i] != prev;
        const unsigned mask = __ballot_sync(0xffffffffu, keep);
        const int pos = off + __popc(mask & ((1u << lane) - 1u));
        if (keep && pos < KC) scratch[pos] = m[i];
        off += __popc(mask);
        CDR_DEV_ASSERT(i == 0 || m[i] == CDR_EMPTY_KEY || prev >= m[i]);      // the merged sequence is sorted
    }
    for (int e = (off < KC ? off : KC) + lane; e < KC; e += 32) scratch[e] = CDR_EMPTY_KEY;
    __syncwarp();
#pragma unroll
    for (int i = 0; i < NPL; ++i) a[i] = scratch[i * 32 + lane];
    __syncwarp();
}

// The lane for one query inside one CTA (8 warps).  The query's posting chunks are numbered along every token's list; this
// CTA takes chunks first_chunk + warp, then every chunk_stride-th, in list (= rank) order: the postings that pass the filter
// are COMPACTED (ballot + prefix count, no insertion per posting) into a staged list that is therefore already sorted,
// until KC of them are staged or nothing further down can still enter the warp's running list; the staged list is merged
// into the running one with duplicates dropped.  The 8 running lists are then merged pairwise the same way: warp 0
// returns with the CTA's list in `top`.  Any row of the final answer is among the first `limit` <= KC passing postings of
// its (warp, token) chunk set, so nothing is lost by the cuts.
template <int NPL>
__device__ __forceinline__ void tech_cta_list(const TechParams &p, int q, int first_chunk, int chunk_stride,
                                              uint64_t (&top)[NPL], uint64_t *s_stage, uint64_t *s_scratch)
{
    constexpr int KC = NPL * 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntok = min(p.q_ntok[q], p.max_tokens);
    const int32_t *toks = p.q_tokens + (size_t)q * p.max_tokens;
    uint64_t *stage = s_stage + warp * KC, *scratch = s_scratch + warp * KC;
#pragma unroll
    for (int i = 0; i < NPL; ++i) top[i] = CDR_EMPTY_KEY;

    for (int t = 0; t < ntok; ++t) {
        const int32_t tok = toks[t];
        if (tok < 0 || tok >= p.n_index_tokens) continue;            // unknown token: no postings
        const int64_t b0 = p.offsets[tok], b1 = p.offsets[tok + 1];
        int cnt = 0;
        for (int64_t base = b0 + (int64_t)(first_chunk + warp) * 32; base < b1 && cnt < KC; base += (int64_t)chunk_stride * 32) {
            // ranks ascend along the list: once the running list is full and even the best possible key of this chunk
            // (its first rank, any row) is not above its last key, nothing from here on can enter
            const uint64_t tau = __shfl_sync(0xffffffffu, top[NPL - 1], 31);
            if (tau != CDR_EMPTY_KEY) {
                const uint64_t best = ((uint64_t)(0xFFFFFFFFu - p.post_rank[base]) << 32) | 0xFFFFFFFFull;
                if (best <= tau) break;
            }
            const int64_t i = base + lane;
            uint64_t key = CDR_EMPTY_KEY;
            if (i < b1) {
                const uint32_t row = p.post_rows[i];
                const uint32_t rank = p.post_rank[i];
                bool ok = true;
                if (p.call_bitmap) {
                    const int32_t slot = p.call_slot[row];
                    ok = slot >= 0 && slot < p.n_call_slots && ((p.call_bitmap[slot >> 5] >> (slot & 31)) & 1u);
                }
                if (ok && (p.has_from | p.has_to)) {
                    const int64_t ts = p.started_at[row];
                    if (p.has_from && ts < p.date_from) ok = false;
                    if (p.has_to && ts > p.date_to) ok = false;
                }
                if (ok && p.has_tags) ok = (p.tag_bits[row] & p.tag_any) != 0ull;
                if (ok) key = ((uint64_t)(0xFFFFFFFFu - rank) << 32) | (uint64_t)(0xFFFFFFFFu - row);
            }
            const unsigned mask = __ballot_sync(0xffffffffu, key != CDR_EMPTY_KEY);
            const int pos = cnt + __popc(mask & ((1u << lane) - 1u));
            if (key != CDR_EMPTY_KEY && pos < KC) stage[pos] = key;
            cnt += __popc(mask);
        }
        if (cnt == 0) continue;                                       // (warp-uniform)
        for (int e = (cnt < KC ? cnt : KC) + lane; e < KC; e += 32) stage[e] = CDR_EMPTY_KEY;
        __syncwarp();
        warp_merge_distinct<NPL>(top, stage, scratch, lane);
    }

    // the 8 running lists, pairwise
#pragma unroll
    for (int step = 1; step < kTechWarps; step <<= 1) {
        if ((warp & (2 * step - 1)) == step) {
#pragma unroll
            for (int i = 0; i < NPL; ++i) stage[i * 32 + lane] = top[i];
        }
        __syncthreads();
        if ((warp & (2 * step - 1)) == 0) warp_merge_distinct<NPL>(top, s_stage + (warp + step) * KC, scratch, lane);
    }
}

// warp 0: the first `limit` rows of the query's list
template <int NPL>
__device__ __forceinline__ void tech_write_out(const TechParams &p, int q, const uint64_t (&top)[NPL], int lane)
{
    int n = 0;
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
        const int e = i * 32 + lane;
        const bool keep = top[i] != CDR_EMPTY_KEY;
        if (keep && e < p.limit) p.out_ids[(size_t)q * p.limit + e] = p.ids[cdr_key_row(top[i])];
        n += __popc(__ballot_sync(0xffffffffu, keep));
    }
    if (n > p.limit) n = p.limit;
    for (int e = n + lane; e < p.limit; e += 32) p.out_ids[(size_t)q * p.limit + e] = -1;
    if (lane == 0) p.out_n[q] = n;
}

// Unfiltered queries: one CTA per query (2 chunks per warp and token are read, whatever the list length).
template <int NPL>
__global__ void __launch_bounds__(kTechWarps * 32) tech_lane_kernel(const TechParams p)
{
    constexpr int KC = NPL * 32;
    __shared__ uint64_t s_stage[kTechWarps * KC];        // a token's passing postings / the list handed to the merge tree
    __shared__ uint64_t s_scratch[kTechWarps * KC];
    uint64_t top[NPL];
    tech_cta_list<NPL>(p, blockIdx.x, 0, kTechWarps, top, s_stage, s_scratch);
    if (threadIdx.x < 32) tech_write_out<NPL>(p, blockIdx.x, top, threadIdx.x);
}

// Filtered queries: under a selective filter a frequent token's whole list is walked before `limit` rows pass (0.87 ms in
// one CTA for a 0.2 % filter, profiles/r02/README.md 8), so a cluster of 8 CTAs shares the chunks of a query; their 8
// lists go to CTA 0 through distributed shared memory and are merged by the same pairwise tree.
constexpr int kTechCluster = 8;
template <int NPL>
__global__ void __cluster_dims__(kTechCluster, 1, 1) __launch_bounds__(kTechWarps * 32) tech_lane_cluster_kernel(const TechParams p)
{
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    constexpr int KC = NPL * 32;
    __shared__ uint64_t s_stage[kTechWarps * KC];
    __shared__ uint64_t s_scratch[kTechWarps * KC];
    __shared__ uint64_t s_gather[kTechCluster * KC];      // CTA 0: the 8 CTAs' lists
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned crank = cluster.block_rank();
    const int q = blockIdx.x / kTechCluster;
    uint64_t top[NPL];
    tech_cta_list<NPL>(p, q, (int)crank * kTechWarps, kTechCluster * kTechWarps, top, s_stage, s_scratch);
    if (warp == 0) {
        uint64_t *g0 = cluster.map_shared_rank(s_gather, 0);
#pragma unroll
        for (int i = 0; i < NPL; ++i) g0[crank * KC + i * 32 + lane] = top[i];
    }
    cluster.sync();
    if (crank != 0) return;
#pragma unroll
    for (int i = 0; i < NPL; ++i) top[i] = s_gather[warp * KC + i * 32 + lane];
#pragma unroll
    for (int step = 1; step < kTechWarps; step <<= 1) {
        if ((warp & (2 * step - 1)) == step) {
#pragma unroll
            for (int i = 0; i < NPL; ++i) s_stage[warp * KC + i * 32 + lane] = top[i];
        }
        __syncthreads();
        if ((warp & (2 * step - 1)) == 0)
            warp_merge_distinct<NPL>(top, s_stage + (warp + step) * KC, s_scratch + warp * KC, lane);
    }
    if (warp == 0) tech_write_out<NPL>(p, q, top, lane);
}

}  // namespace

extern "C" int32_t cdr_tech_index_create(cdr_tech_index **out, cdr_store *s, const int64_t *post_offsets_host,
                                         int32_t n_tokens, const uint32_t *post_rows_host,
                                         const uint32_t *rank_host)
{
    CDR_REQUIRE(out && s && post_offsets_host && rank_host && n_tokens >= 0, CDR_ERR_INVALID,
                "cdr_tech_index_create: NULL argument");
    CDR_REQUIRE(s->finalized, CDR_ERR_STATE, "cdr_tech_index_create: store not finalized");
    *out = nullptr;
    const int64_t total = post_offsets_host[n_tokens];
    CDR_REQUIRE(total >= 0 && (total == 0 || post_rows_host), CDR_ERR_INVALID, "cdr_tech_index_create: bad postings");
    for (int64_t i = 0; i < total; ++i)
        CDR_REQUIRE((int64_t)post_rows_host[i] < s->n_rows, CDR_ERR_INVALID, "cdr_tech_index_create: posting row out of range");
    // every token's postings in the lane's order (rank ascending = call_started_at DESC, id ASC), rank beside the row
    std::vector<uint32_t> rows_sorted((size_t)(total > 0 ? total : 1)), rank_sorted((size_t)(total > 0 ? total : 1));
    {
        std::vector<uint64_t> keyed;
        for (int32_t t = 0; t < n_tokens; ++t) {
            const int64_t b0 = post_offsets_host[t], b1 = post_offsets_host[t + 1];
            CDR_REQUIRE(b0 >= 0 && b1 >= b0 && b1 <= total, CDR_ERR_INVALID, "cdr_tech_index_create: offsets not monotone");
            keyed.resize((size_t)(b1 - b0));
            for (int64_t i = b0; i < b1; ++i)
                keyed[(size_t)(i - b0)] = ((uint64_t)rank_host[post_rows_host[i]] << 32) | post_rows_host[i];
            std::sort(keyed.begin(), keyed.end());
            for (int64_t i = b0; i < b1; ++i) {
                rows_sorted[(size_t)i] = (uint32_t)(keyed[(size_t)(i - b0)] & 0xFFFFFFFFull);
                rank_sorted[(size_t)i] = (uint32_t)(keyed[(size_t)(i - b0)] >> 32);
            }
        }
    }
    DeviceGuard g(s->device);
    cdr_tech_index *ix = new cdr_tech_index();
    ix->store = s;
    ix->n_tokens = n_tokens;
    ix->n_postings = total;
    cudaError_t e = cudaMalloc(&ix->offsets, (size_t)(n_tokens + 1) * 8);
    if (e == cudaSuccess) e = cudaMalloc(&ix->rows, (size_t)(total > 0 ? total : 1) * 4);
    if (e == cudaSuccess) e = cudaMalloc(&ix->prank, (size_t)(total > 0 ? total : 1) * 4);
    if (e == cudaSuccess) e = cudaMemcpy(ix->offsets, post_offsets_host, (size_t)(n_tokens + 1) * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && total > 0) e = cudaMemcpy(ix->rows, rows_sorted.data(), (size_t)total * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && total > 0) e = cudaMemcpy(ix->prank, rank_sorted.data(), (size_t)total * 4, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        cdr_set_error("cdr_tech_index_create: %s", cudaGetErrorString(e));
        cudaFree(ix->offsets); cudaFree(ix->rows); cudaFree(ix->prank);
        delete ix;
        return e == cudaErrorMemoryAllocation ? CDR_ERR_OOM : CDR_ERR_CUDA;
    }
    *out = ix;
    return CDR_OK;
}

extern "C" int32_t cdr_tech_index_destroy(cdr_tech_index *ix)
{
    if (!ix) return CDR_OK;
    DeviceGuard g(ix->store->device);
    cudaDeviceSynchronize();
    cudaFree(ix->offsets);
    cudaFree(ix->rows);
    cudaFree(ix->prank);
    delete ix;
    return CDR_OK;
}

// Enqueue the lane kernel on `st`; every pointer is a device pointer (d_bm nullable).
int cdr_tech_lane_launch(cdr_tech_index *ix, const int32_t *d_tok, const int32_t *d_nt, int nq, int max_tokens,
                         const uint32_t *d_bm, int64_t n_call_slots, int has_date_from, int64_t date_from_us,
                         int has_date_to, int64_t date_to_us, int has_tag_filter, uint64_t tag_any, int limit,
                         int64_t *d_oid, int32_t *d_on, cudaStream_t st)
{
    cdr_store *s = ix->store;
    TechParams p;
    p.offsets = ix->offsets; p.post_rows = ix->rows; p.post_rank = ix->prank;
    p.ids = s->ids; p.call_slot = s->call_slot; p.started_at = s->started_at; p.tag_bits = s->tag_bits;
    p.call_bitmap = d_bm; p.n_call_slots = n_call_slots; p.n_rows = s->n_rows; p.n_index_tokens = ix->n_tokens;
    p.has_from = has_date_from != 0; p.has_to = has_date_to != 0; p.has_tags = has_tag_filter != 0;
    p.date_from = date_from_us; p.date_to = date_to_us; p.tag_any = tag_any;
    p.q_tokens = d_tok; p.q_ntok = d_nt; p.max_tokens = max_tokens; p.limit = limit;
    p.out_ids = d_oid; p.out_n = d_on;
    // CADENCE_TECH_CLUSTER=0 keeps one CTA per query for filtered requests too (A/B aid)
    static const bool use_cluster = [] { const char *e = getenv("CADENCE_TECH_CLUSTER"); return !(e && e[0] == '0'); }();
    const bool filtered = d_bm != nullptr || p.has_from || p.has_to || p.has_tags;
    if (filtered && use_cluster && limit <= 64) tech_lane_cluster_kernel<2><<<nq * kTechCluster, kTechWarps * 32, 0, st>>>(p);
    else if (limit <= 64) tech_lane_kernel<2><<<nq, kTechWarps * 32, 0, st>>>(p);
    else tech_lane_kernel<8><<<nq, kTechWarps * 32, 0, st>>>(p);
    CDR_LAUNCH_CHECK();
    return CDR_OK;
}

cdr_store *cdr_tech_index_store(cdr_tech_index *ix) { return ix->store; }

extern "C" int32_t cdr_tech_lane_host(cdr_tech_index *ix, const int32_t *token_ids_host, const int32_t *n_tokens_host,
                                      int32_t nq, int32_t max_tokens, const uint32_t *call_slot_bitmap_host,
                                      int64_t n_call_slots, int32_t has_date_from, int64_t date_from_us,
                                      int32_t has_date_to, int64_t date_to_us, int32_t has_tag_filter,
                                      uint64_t tag_any, int32_t limit, int64_t *out_ids_host, int32_t *out_n_host,
                                      void *stream)
{
    CDR_REQUIRE(ix && token_ids_host && n_tokens_host && out_ids_host && out_n_host, CDR_ERR_INVALID,
                "cdr_tech_lane_host: NULL argument");
    CDR_REQUIRE(nq >= 0 && max_tokens >= 1 && max_tokens <= kTechMaxTokens, CDR_ERR_INVALID,
                "cdr_tech_lane_host: need 1 <= max_tokens <= %d", kTechMaxTokens);
    CDR_REQUIRE(limit >= 1 && limit <= CDR_MAX_K, CDR_ERR_UNSUPPORTED, "cdr_tech_lane_host: limit=%d outside [1,%d]",
                limit, CDR_MAX_K);
    if (nq == 0) return CDR_OK;
    cdr_store *s = ix->store;
    DeviceGuard g(s->device);
    cudaStream_t st = (cudaStream_t)stream;
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t b_tok = up((size_t)nq * max_tokens * 4), b_nt = up((size_t)nq * 4);
    const size_t b_bm = call_slot_bitmap_host ? up((size_t)((n_call_slots + 31) / 32 + 1) * 4) : 0;
    const size_t b_oid = up((size_t)nq * limit * 8), b_on = up((size_t)nq * 4);
    unsigned char *buf = (unsigned char *)cdr_thread_device(s->device, b_tok + b_nt + b_bm + b_oid + b_on);   // this thread's staging
    if (!buf) return CDR_ERR_OOM;
    unsigned char *c = buf;
    int32_t *d_tok = (int32_t *)c; c += b_tok;
    int32_t *d_nt = (int32_t *)c; c += b_nt;
    uint32_t *d_bm = call_slot_bitmap_host ? (uint32_t *)c : nullptr; c += b_bm;
    int64_t *d_oid = (int64_t *)c; c += b_oid;
    int32_t *d_on = (int32_t *)c;
    CDR_CUDA(cudaMemcpyAsync(d_tok, token_ids_host, (size_t)nq * max_tokens * 4, cudaMemcpyHostToDevice, st));
    CDR_CUDA(cudaMemcpyAsync(d_nt, n_tokens_host, (size_t)nq * 4, cudaMemcpyHostToDevice, st));
    if (d_bm) {
        CDR_CUDA(cudaMemsetAsync(d_bm, 0, b_bm, st));
        if (n_call_slots > 0)
            CDR_CUDA(cudaMemcpyAsync(d_bm, call_slot_bitmap_host, (size_t)((n_call_slots + 31) / 32) * 4,
                                     cudaMemcpyHostToDevice, st));
    }
    int rc = cdr_tech_lane_launch(ix, d_tok, d_nt, nq, max_tokens, d_bm, n_call_slots, has_date_from, date_from_us,
                                  has_date_to, date_to_us, has_tag_filter, tag_any, limit, d_oid, d_on, st);
    if (rc != CDR_OK) return rc;
    CDR_CUDA(cudaMemcpyAsync(out_ids_host, d_oid, (size_t)nq * limit * 8, cudaMemcpyDeviceToHost, st));
    CDR_CUDA(cudaMemcpyAsync(out_n_host, d_on, (size_t)nq * 4, cudaMemcpyDeviceToHost, st));
    CDR_CUDA(cudaStreamSynchronize(st));
    return CDR_OK;
}
