// store.cu -- resident corpus store (one per reference table) and the on-device synthetic
// corpus generator.
//
// Layout in HBM (reference columns in parentheses):
//   emb_f32   f32  [N, dim] row-major          (chunks.embedding vector(1024), alembic 0001:87)
//   emb_bf16  bf16 [N, dim] row-major, each row L2-normalised then RN-even rounded (derived)
//   inv_norm  f32  [N]  1/||x||                 (pgvector recomputes ||x||^2 per row per query)
//   ids       i64  [N]                          (chunk_id BIGSERIAL, alembic 0001:78)
//   call_slot i32  [N]  dictionary code         (call_id UUID, alembic 0001:79)
//   started_at i64 [N]  microseconds            (call_started_at, alembic 0001:81)
//   tag_bits  u64  [N]  per-call tag mask       (calls.tags TEXT[], alembic 0001:45)
//   valid     1 bit/row                         (embedding IS NOT NULL, app/retrieve.py:318,347)
#include "common.cuh"

#include <climits>
#include <cstring>

namespace {

constexpr int kPadRows = 64;   // inv_norm is over-allocated so tail tiles can copy a full tile

// ---- Philox-4x32-10 (same spec as oracle/synth_ref.c; written independently for the device)
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}
__device__ __forceinline__ int ih4(uint32_t w)
{
    return (int)((w & 0xFFu) + ((w >> 8) & 0xFFu) + ((w >> 16) & 0xFFu) + (w >> 24)) - 510;
}

// One warp per row.  Element block b (4 consecutive elements) is lane-strided: b = j*32 + lane.
// Writes fp32 and/or bf16 (normalised) rows, inv_norm, and per-row metadata.
template <int MAXJ>
__global__ void synth_rows_kernel(float *out_f32, __nv_bfloat16 *out_bf16, float *out_inv_norm,
                                  int64_t *out_ids, int32_t *out_call_slot, int64_t *out_started,
                                  uint64_t *out_tags, uint64_t seed, int64_t first_row, int64_t n,
                                  int dim, int64_t id_base, int rows_per_call, int64_t t0_us,
                                  int64_t period_us)
{
    const int lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= n) return;
    const int64_t grow = first_row + i;
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    const int nblk = dim >> 2;
    int4 s[MAXJ];
    long long sumsq = 0;
#pragma unroll
    for (int j = 0; j < MAXJ; ++j) {
        const int b = j * 32 + lane;
        if (b < nblk) {
            const uint4 w = philox4x32_10(
                make_uint4((uint32_t)(uint64_t)grow, (uint32_t)((uint64_t)grow >> 32), (uint32_t)b, 0u), key);
            s[j] = make_int4(ih4(w.x), ih4(w.y), ih4(w.z), ih4(w.w));
            sumsq += (long long)s[j].x * s[j].x + (long long)s[j].y * s[j].y +
                     (long long)s[j].z * s[j].z + (long long)s[j].w * s[j].w;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sumsq += __shfl_xor_sync(0xffffffffu, sumsq, o);
    float inv = 0.f;
    if (sumsq > 0) inv = __fdiv_rn(1.0f, __fsqrt_rn(__ll2float_rn(sumsq)));

    float nrm2 = 0.f;   // ||x||^2 of the stored fp32 row (for inv_norm)
#pragma unroll
    for (int j = 0; j < MAXJ; ++j) {
        const int b = j * 32 + lane;
        if (b < nblk) {
            float4 v;
            v.x = __fmul_rn(__int2float_rn(s[j].x), inv);
            v.y = __fmul_rn(__int2float_rn(s[j].y), inv);
            v.z = __fmul_rn(__int2float_rn(s[j].z), inv);
            v.w = __fmul_rn(__int2float_rn(s[j].w), inv);
            nrm2 = fmaf(v.x, v.x, nrm2);
            nrm2 = fmaf(v.y, v.y, nrm2);
            nrm2 = fmaf(v.z, v.z, nrm2);
            nrm2 = fmaf(v.w, v.w, nrm2);
            if (out_f32) reinterpret_cast<float4 *>(out_f32 + i * dim)[b] = v;
            s[j] = make_int4(__float_as_int(v.x), __float_as_int(v.y), __float_as_int(v.z),
                             __float_as_int(v.w));
        }
    }
    nrm2 = warp_sum_f32(nrm2);
    const float inv_norm = __fdiv_rn(1.0f, __fsqrt_rn(nrm2));
    if (out_bf16) {
#pragma unroll
        for (int j = 0; j < MAXJ; ++j) {
            const int b = j * 32 + lane;
            if (b < nblk) {
                const __nv_bfloat162 lo = __floats2bfloat162_rn(__int_as_float(s[j].x) * inv_norm,
                                                                __int_as_float(s[j].y) * inv_norm);
                const __nv_bfloat162 hi = __floats2bfloat162_rn(__int_as_float(s[j].z) * inv_norm,
                                                                __int_as_float(s[j].w) * inv_norm);
                uint2 pk;
                pk.x = *reinterpret_cast<const uint32_t *>(&lo);
                pk.y = *reinterpret_cast<const uint32_t *>(&hi);
                reinterpret_cast<uint2 *>(out_bf16 + i * dim)[b] = pk;
            }
        }
    }
    if (lane == 0) {
        if (out_inv_norm) out_inv_norm[i] = inv_norm;
        if (out_ids) {
            out_ids[i] = id_base + grow;
            const int64_t slot = grow / rows_per_call;
            out_call_slot[i] = (int32_t)slot;
            out_started[i] = t0_us + slot * period_us;
            const uint4 w = philox4x32_10(
                make_uint4((uint32_t)(uint64_t)slot, (uint32_t)((uint64_t)slot >> 32), 0u, 1u), key);
            out_tags[i] = (1ull << (w.x & 15u)) | (1ull << (w.y & 15u));
        }
    }
}

// inv_norm + normalised bf16 copy of caller-provided fp32 rows (one warp per row).
__global__ void ingest_rows_kernel(const float *rows, __nv_bfloat16 *out_bf16, float *out_inv_norm,
                                   int64_t n, int dim)
{
    const int lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= n) return;
    const float4 *r = reinterpret_cast<const float4 *>(rows + i * dim);
    const int nblk = dim >> 2;
    float nrm2 = 0.f;
    for (int b = lane; b < nblk; b += 32) {
        const float4 v = r[b];
        nrm2 = fmaf(v.x, v.x, nrm2);
        nrm2 = fmaf(v.y, v.y, nrm2);
        nrm2 = fmaf(v.z, v.z, nrm2);
        nrm2 = fmaf(v.w, v.w, nrm2);
    }
    nrm2 = warp_sum_f32(nrm2);
    const float inv_norm = __fdiv_rn(1.0f, __fsqrt_rn(nrm2));
    if (lane == 0) out_inv_norm[i] = inv_norm;
    if (out_bf16) {
        for (int b = lane; b < nblk; b += 32) {
            const float4 v = r[b];
            const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x * inv_norm, v.y * inv_norm);
            const __nv_bfloat162 hi = __floats2bfloat162_rn(v.z * inv_norm, v.w * inv_norm);
            uint2 pk;
            pk.x = *reinterpret_cast<const uint32_t *>(&lo);
            pk.y = *reinterpret_cast<const uint32_t *>(&hi);
            reinterpret_cast<uint2 *>(out_bf16 + i * dim)[b] = pk;
        }
    }
}

// valid bitmap words for rows [row0, row0+n): valid_u8 == nullptr => all valid.
// row0 must be a multiple of 32 OR the leading partial word is merged with atomicOr.
__global__ void valid_bits_kernel(const uint8_t *valid_u8, uint32_t *bitmap, int64_t row0, int64_t n,
                                  unsigned long long *n_valid)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool in = i < n;
    const bool v = in && (valid_u8 == nullptr || valid_u8[i] != 0);
    const int64_t row = row0 + i;
    if (v) atomicOr(&bitmap[row >> 5], 1u << (row & 31));
    const unsigned b = __ballot_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(n_valid, (unsigned long long)__popc(b));
}

__global__ void expand_valid_kernel(const uint32_t *bitmap, int64_t row0, int64_t n, uint8_t *out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const int64_t r = row0 + i;
        out[i] = (bitmap[r >> 5] >> (r & 31)) & 1u;
    }
}

// flag[0] |= 1 when ids are not strictly increasing
__global__ void check_ids_kernel(const int64_t *ids, int64_t n, unsigned long long *flag)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i + 1 < n && ids[i] >= ids[i + 1]) atomicOr(flag, 1ull);
}

// ids[i] -> row position (binary search over the strictly increasing id column), -1 when absent
__global__ void lookup_ids_kernel(const int64_t *store_ids, int64_t n_rows, const int64_t *ids, int64_t n,
                                  int64_t *out_rows, unsigned long long *missing)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t want = ids[i];
    int64_t lo = 0, hi = n_rows;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (store_ids[mid] < want) lo = mid + 1;
        else hi = mid;
    }
    const bool found = lo < n_rows && store_ids[lo] == want;
    out_rows[i] = found ? lo : -1;
    if (!found) atomicAdd(missing, 1ull);
}

// UPDATE ... SET embedding = :e WHERE id = :id for n rows (one warp per row): the fp32 row, its inverse
// norm, its normalised bf16 copy and the `embedding IS NOT NULL` bit.
__global__ void update_rows_kernel(const float *src, const int64_t *rows, int64_t n, int dim, float *emb_f32,
                                   __nv_bfloat16 *emb_bf16, float *inv_norm, uint32_t *valid,
                                   unsigned long long *newly_valid)
{
    const int lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= n) return;
    const int64_t row = rows[i];
    const float4 *r = reinterpret_cast<const float4 *>(src + i * dim);
    const int nblk = dim >> 2;
    float nrm2 = 0.f;
    for (int b = lane; b < nblk; b += 32) {
        const float4 v = r[b];
        nrm2 = fmaf(v.x, v.x, nrm2);
        nrm2 = fmaf(v.y, v.y, nrm2);
        nrm2 = fmaf(v.z, v.z, nrm2);
        nrm2 = fmaf(v.w, v.w, nrm2);
    }
    nrm2 = warp_sum_f32(nrm2);
    const float inv = __fdiv_rn(1.0f, __fsqrt_rn(nrm2));
    for (int b = lane; b < nblk; b += 32) {
        const float4 v = r[b];
        if (emb_f32) reinterpret_cast<float4 *>(emb_f32 + row * dim)[b] = v;
        if (emb_bf16) {
            const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x * inv, v.y * inv);
            const __nv_bfloat162 hi = __floats2bfloat162_rn(v.z * inv, v.w * inv);
            uint2 pk;
            pk.x = *reinterpret_cast<const uint32_t *>(&lo);
            pk.y = *reinterpret_cast<const uint32_t *>(&hi);
            reinterpret_cast<uint2 *>(emb_bf16 + row * dim)[b] = pk;
        }
    }
    if (lane == 0) {
        inv_norm[row] = inv;
        const uint32_t bit = 1u << (row & 31);
        const uint32_t old = atomicOr(&valid[row >> 5], bit);
        if (!(old & bit)) atomicAdd(newly_valid, 1ull);
    }
}

int alloc_store(cdr_store *s)
{
    const int64_t cap = s->capacity;
    const size_t d = (size_t)s->dim;
    if (s->flags & CDR_STORE_FP32) CDR_CUDA(cudaMalloc(&s->emb_f32, (size_t)cap * d * 4));
    if (s->flags & CDR_STORE_BF16) CDR_CUDA(cudaMalloc(&s->emb_bf16, (size_t)cap * d * 2));
    CDR_CUDA(cudaMalloc(&s->inv_norm, (size_t)(cap + kPadRows) * 4));
    CDR_CUDA(cudaMemset(s->inv_norm, 0, (size_t)(cap + kPadRows) * 4));
    CDR_CUDA(cudaMalloc(&s->ids, (size_t)cap * 8));
    CDR_CUDA(cudaMalloc(&s->call_slot, (size_t)cap * 4));
    CDR_CUDA(cudaMalloc(&s->started_at, (size_t)cap * 8));
    CDR_CUDA(cudaMalloc(&s->tag_bits, (size_t)cap * 8));
    const size_t words = (size_t)((cap + 31) / 32) + 4;
    CDR_CUDA(cudaMalloc(&s->valid, words * 4));
    CDR_CUDA(cudaMemset(s->valid, 0, words * 4));
    CDR_CUDA(cudaMalloc(&s->d_scratch, 64 * sizeof(unsigned long long)));
    CDR_CUDA(cudaMemset(s->d_scratch, 0, 64 * sizeof(unsigned long long)));
    return CDR_OK;
}

void free_ws(ScanWorkspace &w)
{
    cudaFree(w.cta_keys);
    cudaFree(w.tile_ctr);
    cudaFree(w.row_list);
    cudaFree(w.gemm_ws);
    w = ScanWorkspace();
}

}  // namespace

int cdr_ws_reserve(void **ptr, size_t *have, size_t need)
{
    if (*have >= need && *ptr != nullptr) return CDR_OK;
    // Growing a workspace: the old buffer may still be in use by work queued on the stream,
    // so synchronise the device before releasing it (rare: only on growth).
    if (*ptr) {
        cudaDeviceSynchronize();
        cudaFree(*ptr);
        *ptr = nullptr;
        *have = 0;
    }
    size_t sz = need < 4096 ? 4096 : need;
    cudaError_t e = cudaMalloc(ptr, sz);
    if (e != cudaSuccess) {
        cdr_set_error("workspace cudaMalloc(%zu) failed: %s", sz, cudaGetErrorString(e));
        *ptr = nullptr;
        return CDR_ERR_OOM;
    }
    *have = sz;
    return CDR_OK;
}

extern "C" int32_t cdr_store_create(cdr_store **out, int32_t device, int64_t capacity_rows,
                                    int32_t dim, uint32_t flags)
{
    CDR_REQUIRE(out != nullptr, CDR_ERR_INVALID, "cdr_store_create: out is NULL");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        cdr_set_error("cdr_store_create: no CUDA device visible; this engine has no CPU fallback");
        return CDR_ERR_NO_DEVICE;
    }
    CDR_REQUIRE(device >= 0 && device < ndev, CDR_ERR_INVALID, "cdr_store_create: device %d out of range [0,%d)", device, ndev);
    CDR_REQUIRE(capacity_rows > 0 && capacity_rows < 0xFFFFFFF0ll, CDR_ERR_INVALID,
                "cdr_store_create: capacity_rows %lld out of range", (long long)capacity_rows);
    CDR_REQUIRE(dim >= 128 && dim <= 4096 && dim % 128 == 0, CDR_ERR_UNSUPPORTED,
                "cdr_store_create: dim %d must be a multiple of 128 in [128,4096]", dim);
    CDR_REQUIRE((flags & (CDR_STORE_FP32 | CDR_STORE_BF16)) != 0 &&
                    (flags & ~(CDR_STORE_FP32 | CDR_STORE_BF16)) == 0,
                CDR_ERR_INVALID, "cdr_store_create: flags must be FP32 and/or BF16");
    DeviceGuard g(device);
    cdr_store *s = new cdr_store();
    s->device = device;
    s->dim = dim;
    s->flags = flags;
    s->capacity = capacity_rows;
    cudaDeviceGetAttribute(&s->sm_count, cudaDevAttrMultiProcessorCount, device);
    int rc = alloc_store(s);
    if (rc != CDR_OK) {
        cdr_store_destroy(s);
        return rc;
    }
    *out = s;
    return CDR_OK;
}

extern "C" int32_t cdr_store_destroy(cdr_store *s)
{
    if (!s) return CDR_OK;
    DeviceGuard g(s->device);
    cudaDeviceSynchronize();
    cudaFree(s->emb_f32);
    cudaFree(s->emb_bf16);
    cudaFree(s->inv_norm);
    cudaFree(s->ids);
    cudaFree(s->call_slot);
    cudaFree(s->started_at);
    cudaFree(s->tag_bits);
    cudaFree(s->valid);
    cudaFree(s->d_scratch);
    for (auto &kv : s->ws) free_ws(kv.second);
    for (auto &m : s->ws_pipe)
        for (auto &kv : m) free_ws(kv.second);
    delete s;
    return CDR_OK;
}

extern "C" int32_t cdr_store_append(cdr_store *s, const float *rows_f32, const int64_t *ids,
                                    const int32_t *call_slot, const int64_t *started_at_us,
                                    const uint64_t *tag_bits, const uint8_t *valid_u8, int64_t n,
                                    int32_t is_device, void *stream)
{
    CDR_REQUIRE(s != nullptr, CDR_ERR_INVALID, "cdr_store_append: store is NULL");
    CDR_REQUIRE(n >= 0 && rows_f32 && ids, CDR_ERR_INVALID, "cdr_store_append: rows/ids required");
    if (n == 0) return CDR_OK;
    DeviceGuard g(s->device);
    std::lock_guard<std::mutex> lk(s->mu);
    // (checked under the lock: two concurrent appends must not both pass against the same row count)
    CDR_REQUIRE(s->n_rows + n <= s->capacity, CDR_ERR_OOM,
                "cdr_store_append: %lld + %lld rows exceed capacity %lld", (long long)s->n_rows,
                (long long)n, (long long)s->capacity);
    cudaStream_t st = (cudaStream_t)stream;
    // A sealed store keeps serving while it grows (new chunks of an ingested call): the id order is then
    // verified BEFORE anything is written, so a rejected batch leaves the store untouched.
    const bool sealed = s->finalized;
    int64_t batch_last_id = 0;
    if (sealed) {
        std::vector<int64_t> h_ids((size_t)n);
        if (is_device) {
            CDR_CUDA(cudaMemcpyAsync(h_ids.data(), ids, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
            CDR_CUDA(cudaStreamSynchronize(st));
        } else {
            memcpy(h_ids.data(), ids, (size_t)n * 8);
        }
        int64_t prev = s->n_rows > 0 ? s->last_id : INT64_MIN;
        for (int64_t i = 0; i < n; ++i) {
            CDR_REQUIRE(h_ids[(size_t)i] > prev || (i == 0 && s->n_rows == 0), CDR_ERR_UNSORTED_IDS,
                        "cdr_store_append: id %lld at position %lld is not above the previous id %lld "
                        "(rows must arrive in id order)", (long long)h_ids[(size_t)i], (long long)i, (long long)prev);
            prev = h_ids[(size_t)i];
        }
        batch_last_id = prev;
    }
    const cudaMemcpyKind kind = is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    const int64_t r0 = s->n_rows;
    const size_t d = (size_t)s->dim;

    // fp32 rows land in the store (or in a temporary when the store keeps bf16 only)
    float *dst_rows = nullptr;
    float *tmp_rows = nullptr;
    const int64_t chunk = 1 << 16;
    if (s->flags & CDR_STORE_FP32) {
        dst_rows = s->emb_f32 + (size_t)r0 * d;
        CDR_CUDA(cudaMemcpyAsync(dst_rows, rows_f32, (size_t)n * d * 4, kind, st));
        const int wpb = 8;
        ingest_rows_kernel<<<(unsigned)((n + wpb - 1) / wpb), wpb * 32, 0, st>>>(
            dst_rows, s->emb_bf16 ? s->emb_bf16 + (size_t)r0 * d : nullptr, s->inv_norm + r0, n, s->dim);
        CDR_LAUNCH_CHECK();
    } else {
        CDR_CUDA(cudaMalloc(&tmp_rows, (size_t)(n < chunk ? n : chunk) * d * 4));
        for (int64_t o = 0; o < n; o += chunk) {
            const int64_t m = (n - o) < chunk ? (n - o) : chunk;
            CDR_CUDA(cudaMemcpyAsync(tmp_rows, rows_f32 + (size_t)o * d, (size_t)m * d * 4, kind, st));
            const int wpb = 8;
            ingest_rows_kernel<<<(unsigned)((m + wpb - 1) / wpb), wpb * 32, 0, st>>>(
                tmp_rows, s->emb_bf16 + (size_t)(r0 + o) * d, s->inv_norm + r0 + o, m, s->dim);
            CDR_LAUNCH_CHECK();
        }
        CDR_CUDA(cudaStreamSynchronize(st));
        cudaFree(tmp_rows);
    }
    CDR_CUDA(cudaMemcpyAsync(s->ids + r0, ids, (size_t)n * 8, kind, st));
    if (call_slot) CDR_CUDA(cudaMemcpyAsync(s->call_slot + r0, call_slot, (size_t)n * 4, kind, st));
    else CDR_CUDA(cudaMemsetAsync(s->call_slot + r0, 0, (size_t)n * 4, st));
    if (started_at_us) CDR_CUDA(cudaMemcpyAsync(s->started_at + r0, started_at_us, (size_t)n * 8, kind, st));
    else CDR_CUDA(cudaMemsetAsync(s->started_at + r0, 0, (size_t)n * 8, st));
    if (tag_bits) CDR_CUDA(cudaMemcpyAsync(s->tag_bits + r0, tag_bits, (size_t)n * 8, kind, st));
    else CDR_CUDA(cudaMemsetAsync(s->tag_bits + r0, 0, (size_t)n * 8, st));

    uint8_t *valid_dev = nullptr;
    if (valid_u8 && !is_device) {
        CDR_CUDA(cudaMalloc(&valid_dev, (size_t)n));
        CDR_CUDA(cudaMemcpyAsync(valid_dev, valid_u8, (size_t)n, cudaMemcpyHostToDevice, st));
    }
    valid_bits_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(
        valid_u8 ? (is_device ? valid_u8 : valid_dev) : nullptr, s->valid, r0, n, s->d_scratch + 1);
    CDR_LAUNCH_CHECK();
    // host buffers may be reused by the caller as soon as we return
    CDR_CUDA(cudaStreamSynchronize(st));
    if (valid_dev) cudaFree(valid_dev);
    s->n_rows += n;
    if (sealed) {
        unsigned long long h[2] = {0, 0};
        CDR_CUDA(cudaMemcpy(h, s->d_scratch, sizeof(h), cudaMemcpyDeviceToHost));
        s->n_valid = (int64_t)h[1];
        s->any_invalid = s->n_valid != s->n_rows;
        s->last_id = batch_last_id;
    }
    return CDR_OK;
}

extern "C" int32_t cdr_store_update_embeddings(cdr_store *s, const int64_t *ids_host, const float *rows_f32_host,
                                               int64_t n, void *stream)
{
    CDR_REQUIRE(s != nullptr, CDR_ERR_INVALID, "cdr_store_update_embeddings: store is NULL");
    CDR_REQUIRE(s->finalized, CDR_ERR_STATE, "cdr_store_update_embeddings: store not finalized (ids are located by "
                "binary search over the sealed id column)");
    CDR_REQUIRE(n >= 0 && (n == 0 || (ids_host && rows_f32_host)), CDR_ERR_INVALID,
                "cdr_store_update_embeddings: ids/rows required");
    if (n == 0) return CDR_OK;
    DeviceGuard g(s->device);
    std::lock_guard<std::mutex> lk(s->mu);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t d = (size_t)s->dim;
    // this thread's device staging: [missing / newly-valid counters | ids | rows | source vectors]
    const size_t off_ids = 16, off_rows = off_ids + (size_t)n * 8, off_src = (off_rows + (size_t)n * 8 + 255) & ~(size_t)255;
    unsigned char *stage = (unsigned char *)cdr_thread_device(s->device, off_src + (size_t)n * d * 4);
    if (!stage) return CDR_ERR_OOM;
    unsigned long long *d_cnt = reinterpret_cast<unsigned long long *>(stage);    // [0] missing ids, [1] rows that turned NOT NULL
    int64_t *d_ids = reinterpret_cast<int64_t *>(stage + off_ids);
    int64_t *d_rows = reinterpret_cast<int64_t *>(stage + off_rows);
    float *d_src = reinterpret_cast<float *>(stage + off_src);
    unsigned long long h[2] = {0, 0};
    cudaError_t e = cudaMemsetAsync(d_cnt, 0, 16, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_ids, ids_host, (size_t)n * 8, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        lookup_ids_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(s->ids, s->n_rows, d_ids, n, d_rows, d_cnt);
        g_cdr_launches.fetch_add(1, std::memory_order_relaxed);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(h, d_cnt, 16, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e == cudaSuccess && h[0] != 0) {
        cdr_set_error("cdr_store_update_embeddings: %llu of %lld ids are not in the store; nothing was updated", h[0],
                      (long long)n);
        return CDR_ERR_INVALID;
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_src, rows_f32_host, (size_t)n * d * 4, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        const int wpb = 8;
        update_rows_kernel<<<(unsigned)((n + wpb - 1) / wpb), wpb * 32, 0, st>>>(d_src, d_rows, n, s->dim, s->emb_f32,
                                                                                  s->emb_bf16, s->inv_norm, s->valid,
                                                                                  d_cnt + 1);
        g_cdr_launches.fetch_add(1, std::memory_order_relaxed);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(h, d_cnt, 16, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        cdr_set_error("cdr_store_update_embeddings: %s", cudaGetErrorString(e));
        return CDR_ERR_CUDA;
    }
    s->n_valid += (int64_t)h[1];
    s->any_invalid = s->n_valid != s->n_rows;
    // keep the running count of NOT NULL rows (used when the sealed store grows) in step
    CDR_CUDA(cudaMemcpy(s->d_scratch + 1, &s->n_valid, 8, cudaMemcpyHostToDevice));
    return CDR_OK;
}

template <int MAXJ>
static int launch_synth(float *o32, __nv_bfloat16 *o16, float *oin, int64_t *oid, int32_t *ocs,
                        int64_t *ost, uint64_t *otg, uint64_t seed, int64_t first_row, int64_t n,
                        int dim, int64_t id_base, int rpc, int64_t t0, int64_t period, cudaStream_t st)
{
    const int wpb = 8;
    // grid is limited to 2^31-1 blocks: fine up to 1.7e10 rows
    synth_rows_kernel<MAXJ><<<(unsigned)((n + wpb - 1) / wpb), wpb * 32, 0, st>>>(
        o32, o16, oin, oid, ocs, ost, otg, seed, first_row, n, dim, id_base, rpc, t0, period);
    CDR_LAUNCH_CHECK();
    return CDR_OK;
}

static int synth_dispatch(float *o32, __nv_bfloat16 *o16, float *oin, int64_t *oid, int32_t *ocs,
                          int64_t *ost, uint64_t *otg, uint64_t seed, int64_t first_row, int64_t n,
                          int dim, int64_t id_base, int rpc, int64_t t0, int64_t period, cudaStream_t st)
{
    if (dim <= 1024) return launch_synth<8>(o32, o16, oin, oid, ocs, ost, otg, seed, first_row, n, dim, id_base, rpc, t0, period, st);
    if (dim <= 2048) return launch_synth<16>(o32, o16, oin, oid, ocs, ost, otg, seed, first_row, n, dim, id_base, rpc, t0, period, st);
    return launch_synth<32>(o32, o16, oin, oid, ocs, ost, otg, seed, first_row, n, dim, id_base, rpc, t0, period, st);
}

extern "C" int32_t cdr_store_append_synthetic(cdr_store *s, uint64_t seed, int64_t first_row,
                                              int64_t n, int64_t id_base, int32_t rows_per_call,
                                              int64_t t0_us, int64_t call_period_us, void *stream)
{
    CDR_REQUIRE(s != nullptr, CDR_ERR_INVALID, "cdr_store_append_synthetic: store is NULL");
    CDR_REQUIRE(!s->finalized, CDR_ERR_STATE, "cdr_store_append_synthetic: store already finalized");
    CDR_REQUIRE(n >= 0 && first_row >= 0 && rows_per_call > 0, CDR_ERR_INVALID,
                "cdr_store_append_synthetic: bad n/first_row/rows_per_call");
    if (n == 0) return CDR_OK;
    DeviceGuard g(s->device);
    std::lock_guard<std::mutex> lk(s->mu);
    CDR_REQUIRE(s->n_rows + n <= s->capacity, CDR_ERR_OOM,
                "cdr_store_append_synthetic: %lld + %lld rows exceed capacity %lld",
                (long long)s->n_rows, (long long)n, (long long)s->capacity);
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t r0 = s->n_rows;
    const size_t d = (size_t)s->dim;
    int rc = synth_dispatch(s->emb_f32 ? s->emb_f32 + (size_t)r0 * d : nullptr,
                            s->emb_bf16 ? s->emb_bf16 + (size_t)r0 * d : nullptr, s->inv_norm + r0,
                            s->ids + r0, s->call_slot + r0, s->started_at + r0, s->tag_bits + r0, seed,
                            first_row, n, s->dim, id_base, rows_per_call, t0_us, call_period_us, st);
    if (rc != CDR_OK) return rc;
    valid_bits_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(nullptr, s->valid, r0, n,
                                                                    s->d_scratch + 1);
    CDR_LAUNCH_CHECK();
    s->n_rows += n;
    return CDR_OK;
}

extern "C" int32_t cdr_synth_rows(float *out_dev, uint64_t seed, int64_t first_row, int64_t n,
                                  int32_t dim, void *stream)
{
    CDR_REQUIRE(out_dev && n >= 0 && dim >= 4 && dim % 4 == 0 && dim <= 4096, CDR_ERR_INVALID,
                "cdr_synth_rows: bad arguments");
    if (n == 0) return CDR_OK;
    return synth_dispatch(out_dev, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, seed, first_row,
                          n, dim, 0, 1, 0, 0, (cudaStream_t)stream);
}

extern "C" int32_t cdr_store_finalize(cdr_store *s, void *stream)
{
    CDR_REQUIRE(s != nullptr, CDR_ERR_INVALID, "cdr_store_finalize: store is NULL");
    if (s->finalized) return CDR_OK;
    DeviceGuard g(s->device);
    std::lock_guard<std::mutex> lk(s->mu);
    cudaStream_t st = (cudaStream_t)stream;
    if (s->n_rows > 1) {
        check_ids_kernel<<<(unsigned)((s->n_rows + 255) / 256), 256, 0, st>>>(s->ids, s->n_rows,
                                                                               s->d_scratch);
        CDR_LAUNCH_CHECK();
    }
    unsigned long long h[2] = {0, 0};
    CDR_CUDA(cudaMemcpyAsync(h, s->d_scratch, sizeof(h), cudaMemcpyDeviceToHost, st));
    CDR_CUDA(cudaStreamSynchronize(st));
    CDR_REQUIRE(h[0] == 0, CDR_ERR_UNSORTED_IDS,
                "cdr_store_finalize: ids must be strictly increasing in append order "
                "(load rows ORDER BY chunk_id)");
    s->n_valid = (int64_t)h[1];
    s->any_invalid = s->n_valid != s->n_rows;
    if (s->n_rows > 0) CDR_CUDA(cudaMemcpy(&s->last_id, s->ids + (s->n_rows - 1), 8, cudaMemcpyDeviceToHost));
    s->finalized = true;
    return CDR_OK;
}

extern "C" int32_t cdr_store_info(const cdr_store *s, int64_t *rows, int32_t *dim, uint32_t *flags,
                                  int64_t *n_valid, int32_t *device)
{
    CDR_REQUIRE(s != nullptr, CDR_ERR_INVALID, "cdr_store_info: store is NULL");
    if (rows) *rows = s->n_rows;
    if (dim) *dim = s->dim;
    if (flags) *flags = s->flags;
    if (n_valid) *n_valid = s->n_valid;
    if (device) *device = s->device;
    return CDR_OK;
}

extern "C" int32_t cdr_store_read_rows(cdr_store *s, int64_t first_row, int64_t n,
                                       float *out_f32_host, uint16_t *out_bf16_host,
                                       int64_t *out_ids_host, int32_t *out_call_slot_host,
                                       int64_t *out_started_at_host, uint64_t *out_tag_bits_host,
                                       float *out_inv_norm_host)
{
    CDR_REQUIRE(s != nullptr, CDR_ERR_INVALID, "cdr_store_read_rows: store is NULL");
    CDR_REQUIRE(first_row >= 0 && n >= 0 && first_row + n <= s->n_rows, CDR_ERR_INVALID,
                "cdr_store_read_rows: range [%lld,+%lld) outside [0,%lld)", (long long)first_row,
                (long long)n, (long long)s->n_rows);
    DeviceGuard g(s->device);
    const size_t d = (size_t)s->dim;
    CDR_CUDA(cudaDeviceSynchronize());
    if (out_f32_host) {
        CDR_REQUIRE(s->emb_f32 != nullptr, CDR_ERR_STATE, "cdr_store_read_rows: no fp32 rows resident");
        CDR_CUDA(cudaMemcpy(out_f32_host, s->emb_f32 + (size_t)first_row * d, (size_t)n * d * 4, cudaMemcpyDeviceToHost));
    }
    if (out_bf16_host) {
        CDR_REQUIRE(s->emb_bf16 != nullptr, CDR_ERR_STATE, "cdr_store_read_rows: no bf16 rows resident");
        CDR_CUDA(cudaMemcpy(out_bf16_host, s->emb_bf16 + (size_t)first_row * d, (size_t)n * d * 2, cudaMemcpyDeviceToHost));
    }
    if (out_ids_host) CDR_CUDA(cudaMemcpy(out_ids_host, s->ids + first_row, (size_t)n * 8, cudaMemcpyDeviceToHost));
    if (out_call_slot_host) CDR_CUDA(cudaMemcpy(out_call_slot_host, s->call_slot + first_row, (size_t)n * 4, cudaMemcpyDeviceToHost));
    if (out_started_at_host) CDR_CUDA(cudaMemcpy(out_started_at_host, s->started_at + first_row, (size_t)n * 8, cudaMemcpyDeviceToHost));
    if (out_tag_bits_host) CDR_CUDA(cudaMemcpy(out_tag_bits_host, s->tag_bits + first_row, (size_t)n * 8, cudaMemcpyDeviceToHost));
    if (out_inv_norm_host) CDR_CUDA(cudaMemcpy(out_inv_norm_host, s->inv_norm + first_row, (size_t)n * 4, cudaMemcpyDeviceToHost));
    return CDR_OK;
}

extern "C" int32_t cdr_store_copy_rows_device(cdr_store *s, int64_t first_row, int64_t n, float *out_f32_dev,
                                              uint16_t *out_bf16_dev, void *stream)
{
    CDR_REQUIRE(s != nullptr, CDR_ERR_INVALID, "cdr_store_copy_rows_device: store is NULL");
    CDR_REQUIRE(first_row >= 0 && n >= 0 && first_row + n <= s->n_rows, CDR_ERR_INVALID,
                "cdr_store_copy_rows_device: range [%lld,+%lld) outside [0,%lld)", (long long)first_row,
                (long long)n, (long long)s->n_rows);
    if (n == 0) return CDR_OK;
    DeviceGuard g(s->device);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t d = (size_t)s->dim;
    if (out_f32_dev) {
        CDR_REQUIRE(s->emb_f32 != nullptr, CDR_ERR_STATE, "cdr_store_copy_rows_device: no fp32 rows resident");
        CDR_CUDA(cudaMemcpyAsync(out_f32_dev, s->emb_f32 + (size_t)first_row * d, (size_t)n * d * 4,
                                 cudaMemcpyDeviceToDevice, st));
    }
    if (out_bf16_dev) {
        CDR_REQUIRE(s->emb_bf16 != nullptr, CDR_ERR_STATE, "cdr_store_copy_rows_device: no bf16 rows resident");
        CDR_CUDA(cudaMemcpyAsync(out_bf16_dev, s->emb_bf16 + (size_t)first_row * d, (size_t)n * d * 2,
                                 cudaMemcpyDeviceToDevice, st));
    }
    return CDR_OK;
}

extern "C" int32_t cdr_store_read_valid(cdr_store *s, int64_t first_row, int64_t n, uint8_t *out_valid_u8_host)
{
    CDR_REQUIRE(s != nullptr && out_valid_u8_host != nullptr, CDR_ERR_INVALID, "cdr_store_read_valid: NULL argument");
    CDR_REQUIRE(first_row >= 0 && n >= 0 && first_row + n <= s->n_rows, CDR_ERR_INVALID,
                "cdr_store_read_valid: range [%lld,+%lld) outside [0,%lld)", (long long)first_row, (long long)n,
                (long long)s->n_rows);
    if (n == 0) return CDR_OK;
    DeviceGuard g(s->device);
    uint8_t *tmp = nullptr;
    CDR_CUDA(cudaMalloc(&tmp, (size_t)n));
    expand_valid_kernel<<<(unsigned)((n + 255) / 256), 256>>>(s->valid, first_row, n, tmp);
    CDR_LAUNCH_CHECK();
    cudaError_t e = cudaMemcpy(out_valid_u8_host, tmp, (size_t)n, cudaMemcpyDeviceToHost);
    cudaFree(tmp);
    CDR_CUDA(e);
    return CDR_OK;
}
