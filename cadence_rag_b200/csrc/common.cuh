// common.cuh -- shared host/device helpers for libcadence_dense (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <map>
#include <mutex>
#include <vector>

#include "../../include/cadence_dense.h"

// ------------------------------------------------------------------------------------ errors
void cdr_set_error(const char *fmt, ...);
extern std::atomic<int64_t> g_cdr_launches;

#define CDR_CUDA(expr)                                                                        \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            cdr_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,   \
                          __LINE__);                                                          \
            return _e == cudaErrorMemoryAllocation ? CDR_ERR_OOM : CDR_ERR_CUDA;              \
        }                                                                                     \
    } while (0)

#define CDR_LAUNCH_CHECK()                                                                    \
    do {                                                                                      \
        g_cdr_launches.fetch_add(1, std::memory_order_relaxed);                               \
        CDR_CUDA(cudaGetLastError());                                                         \
    } while (0)

#define CDR_REQUIRE(cond, code, ...)                                                          \
    do {                                                                                      \
        if (!(cond)) {                                                                        \
            cdr_set_error(__VA_ARGS__);                                                       \
            return (code);                                                                    \
        }                                                                                     \
    } while (0)

// ------------------------------------------------------------------------------------ store
struct ScanWorkspace {
    uint64_t *cta_keys = nullptr;   // [nq_cap, grid, KC] packed candidate keys from K1
    size_t cta_keys_bytes = 0;
    uint32_t *row_list = nullptr;       // K1 selective filters: compact list of allowed rows (+ count behind it)
    size_t row_list_bytes = 0;
    unsigned int *tile_ctr = nullptr;   // [nq_cap] K1 work-stealing counters (zero between launches)
    size_t tile_ctr_bytes = 0;
    void *gemm_ws = nullptr;        // K2 candidate lists, thresholds, bf16 queries
    size_t gemm_ws_bytes = 0;
};

// Request / response staging of the calling THREAD for the *_host entry points: a pinned host mirror and a
// device buffer per (thread, device), so concurrent requests never share staging even when they share a
// stream (the kernels' own workspaces above are per stream and protected by stream order + the store lock).
// Both return nullptr (error text set) when the allocation fails; grown on demand -- the thread's previous
// call has synchronised by then -- and kept for the thread's life.
void *cdr_thread_pinned(size_t need);
void *cdr_thread_device(int device, size_t need);

struct cdr_store {
    int device = 0;
    int dim = 0;
    int sm_count = 0;
    uint32_t flags = 0;
    int64_t capacity = 0;
    int64_t n_rows = 0;
    int64_t n_valid = 0;
    bool finalized = false;
    bool any_invalid = false;
    int64_t last_id = INT64_MIN;   // id of the last row (host copy; maintained once the store is sealed)

    float *emb_f32 = nullptr;            // [capacity, dim]
    __nv_bfloat16 *emb_bf16 = nullptr;   // [capacity, dim], rows L2-normalised before rounding
    float *inv_norm = nullptr;           // [capacity + pad] 1/||x|| in fp32 (inf for zero rows)
    int64_t *ids = nullptr;              // [capacity]
    int32_t *call_slot = nullptr;        // [capacity]
    int64_t *started_at = nullptr;       // [capacity] microseconds
    uint64_t *tag_bits = nullptr;        // [capacity]
    uint32_t *valid = nullptr;           // bitmap [ceil(capacity/32)] embedding IS NOT NULL
    unsigned long long *d_scratch = nullptr;  // small device scratch (counters / flags)

    std::mutex mu;
    std::map<cudaStream_t, ScanWorkspace> ws;
    std::map<cudaStream_t, ScanWorkspace> ws_pipe[2];   // the two alternating workspaces of a pipelined sharded step
};

int cdr_ws_reserve(void **ptr, size_t *have, size_t need);

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        // Every entry point builds one of these before it launches anything.  A non-sticky error left behind by an
        // earlier failed runtime call of ANOTHER library in the process (a peer-access probe on a one-GPU box answers
        // "invalid device ordinal") would otherwise be picked up by this call's first cudaGetLastError() and blamed on a
        // kernel launch that was fine; sticky errors (real faults) survive this and still fail the call.
        (void)cudaGetLastError();
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// One rank's view of a peer-memory exchange group (peer.cu) for ONE epoch, passed by value to the kernels that take part in
// the exchange: the K4p kernel, and -- fused -- the finalize kernels of the scan lanes, whose ordering CTA pushes the
// query's final list straight into the peers' buffers, waits for theirs and merges (no local list round trip, one launch
// less per step).  world == 0: no exchange.
constexpr int kPeerFusedRanks = 8;      // the fused form keeps its merge scratch in static shared memory:
constexpr int kPeerFusedK = 64;         // world <= 8 ranks x k <= 64 entries x 16 bytes = 8 KB
struct PeerLink {
    unsigned char *base[CDR_PEER_MAX_RANKS];   // every rank's receive buffer as mapped here (own included)
    int rank = 0, world = 0;
    int max_nq = 0, max_k = 0;
    size_t entry_bytes = 0, flags_off = 0;
    uint32_t epoch = 0;
    int q0 = 0;                                // buffer slot of the launch's first query
};

// K1 + finalize on `st` (exact_scan.cu).  share_reads: score every streamed tile against 3 queries per CTA
// (batches of concurrent exact requests); false = one scan of the corpus per query.  Same bits either way.
// Single-query "ann" lane (exact_scan.cu): candidate pass over the bf16 rows, exact re-score.  Store mutex held by the caller.
// q_index / q_count: the conditional re-run form (see cdr_exact_scan_redo_launch), both null for a plain launch.
// peer (plain launches only): the finalize kernel ends with the exchange of `peer`'s epoch -- out_* then receive the
// MERGED lists of all ranks (see PeerLink).
int cdr_bf16_scan_launch(cdr_store *s, ScanWorkspace &ws, const float *q_dev, int nq, const uint32_t *allow, int k,
                         double *out_score, int64_t *out_id, int32_t *out_n, cudaStream_t st,
                         const int *q_index = nullptr, const int *q_count = nullptr, const PeerLink *peer = nullptr);
// Re-run queries q_index[0 .. min(*q_count, n_slots)) (count on the device) on the exact fp32 lane, results written to
// the queries' own output rows; with a zero count the launches exit at once.
int cdr_exact_scan_redo_launch(cdr_store *s, ScanWorkspace &ws, const float *q_dev, int n_slots, const uint32_t *allow,
                               int k, double *out_score, int64_t *out_id, int32_t *out_n, cudaStream_t st,
                               const int *q_index, const int *q_count);
// K2 (gemm_topk.cu) with the store mutex held by the caller: nq <= 16384, k <= 192, bf16 rows resident.
int cdr_batch_bf16_launch(cdr_store *s, ScanWorkspace &ws, const float *q_dev, int nq, int k, const uint32_t *allow_dev,
                          double *out_score_dev, int64_t *out_id_dev, int32_t *out_n_dev, cudaStream_t st);
// fin: when set, the finalize kernel of the launch goes to another stream -- `scan_done` is recorded on `st` after the scan
// kernels and `fin->stream` waits for it -- so the caller can overlap it (and what it enqueues behind it) with the NEXT
// scan on `st` (peer.cu: the sharded step pipelines chunks of a batch this way).
struct ScanFinalizeOn {
    cudaStream_t stream;
    cudaEvent_t scan_done;
};
int cdr_exact_scan_launch(cdr_store *s, ScanWorkspace &ws, const float *q_dev, int nq, const uint32_t *allow, int k,
                          double *out_score, int64_t *out_id, int32_t *out_n, cudaStream_t st, bool share_reads,
                          const ScanFinalizeOn *fin = nullptr, const PeerLink *peer = nullptr);
bool cdr_scan_peer_fusable(bool bf16_rows, int k, int world);
// The scan lanes behind cdr_search_exact_f32[_shared] / cdr_search_scan_bf16 with an optional fused exchange (abi.cu).
int32_t cdr_search_exact_peer(bool share_reads, cdr_store *s, const float *q_dev, int32_t nq, int32_t k,
                              const uint32_t *allow_dev, double *out_score_dev, int64_t *out_id_dev, int32_t *out_n_dev,
                              void *stream, const PeerLink *peer);
int32_t cdr_search_scan_bf16_peer(cdr_store *s, const float *q_dev, int32_t nq, int32_t k, const uint32_t *allow_dev,
                                  double *out_score_dev, int64_t *out_id_dev, int32_t *out_n_dev, void *stream,
                                  const PeerLink *peer);

// profiling hooks (abi.cu)
void cdr_prof_mark_begin(int kind, cudaStream_t st);
void cdr_prof_mark_end(int kind, cudaStream_t st);
bool cdr_prof_active();

// ------------------------------------------------------------------------------------ device
#ifdef __CUDACC__

#define CDR_EMPTY_KEY 0ull

// Own bounds / protocol checks (compute-sanitizer is not available on every pool): a build with -DCDR_DEBUG_BOUNDS traps
// (launch error, reported through the C ABI) where an index leaves its array or a pipeline invariant breaks.  Compiled
// out of the shipped library.  tools/sanitize_driver.cc runs every kernel family against such a build.
#ifdef CDR_DEBUG_BOUNDS
#define CDR_DEV_ASSERT(cond)                                                                    \
    do {                                                                                        \
        if (!(cond)) {                                                                          \
            printf("CDR_DEV_ASSERT failed: %s (%s:%d) block %d thread %d\n", #cond, __FILE__, __LINE__, (int)blockIdx.x, \
                   (int)threadIdx.x);                                                           \
            __trap();                                                                           \
        }                                                                                       \
    } while (0)
#else
#define CDR_DEV_ASSERT(cond) do { } while (0)
#endif

// Monotone map float -> uint32 (larger float => larger code).  NaN => 1 (below every real
// cosine, above the empty key's 0), so NaN rows sort last but stay eligible (SQL semantics).
__device__ __forceinline__ uint32_t cdr_order_f32(float f)
{
    uint32_t u = __float_as_uint(f);
    if ((u & 0x7FFFFFFFu) > 0x7F800000u) return 1u;
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// key = (order(score) << 32) | (0xFFFFFFFF - row): bigger key = better (score desc, row asc).
__device__ __forceinline__ uint64_t cdr_pack_key(float score, uint32_t row)
{
    return ((uint64_t)cdr_order_f32(score) << 32) | (uint64_t)(0xFFFFFFFFu - row);
}
__device__ __forceinline__ uint32_t cdr_key_row(uint64_t key)
{
    return 0xFFFFFFFFu - (uint32_t)key;
}

// (score desc, NaN last, id asc) "a sorts before b"
__device__ __forceinline__ bool cdr_result_before(double sa, int64_t ia, double sb, int64_t ib)
{
    bool na = sa != sa, nb = sb != sb;
    if (na != nb) return nb;
    if (!na && sa != sb) return sa > sb;
    return ia < ib;
}

// The same order on integers: the B200's fp64 compare path is slow (a 64 x 64 rank count over double scores cost 14.5 us
// in the latency finalize, profiles/r02/latency), so the ranking loops compare order keys.  Monotone map double -> uint64
// (larger score => larger key), NaN => 0 (below every number: NaN last), +-0 => one key (they compare equal).
// cdr_key_before(ka, ia, kb, ib) == cdr_result_before(sa, ia, sb, ib) for ka = cdr_order_f64(sa), kb = cdr_order_f64(sb).
__device__ __forceinline__ uint64_t cdr_order_f64_bits(uint64_t b)
{
    const uint64_t mag = b & 0x7FFFFFFFFFFFFFFFull;
    if (mag > 0x7FF0000000000000ull) return 0ull;
    if (mag == 0ull) return 0x8000000000000000ull;
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ uint64_t cdr_order_f64(double s)
{
    return cdr_order_f64_bits((uint64_t)__double_as_longlong(s));
}
__device__ __forceinline__ bool cdr_key_before(uint64_t ka, int64_t ia, uint64_t kb, int64_t ib)
{
    return ka > kb || (ka == kb && ia < ib);
}

// Number of entries of an ORDERED list (keys / ids, n valid entries, n <= k_cap) that sort before (key, id): branch-free
// binary search with a trip count fixed by k_cap.
__device__ __forceinline__ int cdr_sorted_count_before(const uint64_t *keys, const int64_t *ids, int n, int k_cap, uint64_t key,
                                                       int64_t id)
{
    CDR_DEV_ASSERT(n >= 0 && n <= k_cap);
    int cnt = 0;
    for (int half = 1 << (31 - __clz(k_cap > 1 ? k_cap : 1)); half > 0; half >>= 1) {
        const int pos = cnt + half;
        const bool in = pos <= n;
        const int idx = in ? pos - 1 : 0;
        const bool before = cdr_key_before(keys[idx], ids[idx], key, id);
        cnt = (in && before) ? pos : cnt;
    }
    return cnt;
}

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}

// ---- mbarrier / bulk-copy PTX wrappers
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug traps (launch error) instead of hanging the GPU box.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    if (mbar_try_wait(bar, parity)) return;
    long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000ll) __trap();
    }
}
// The same for waits that are long by design and off the critical path (an epilogue warp waiting for the next
// accumulator): the warp sleeps between polls instead of spinning on the issue port.
__device__ __forceinline__ void mbar_wait_backoff(uint64_t *bar, uint32_t parity, unsigned ns)
{
    if (mbar_try_wait(bar, parity)) return;
    long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (ns) __nanosleep(ns);
        if (clock64() - t0 > 4000000000ll) __trap();
    }
}
// 1-D bulk async copy global -> shared, completion on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes,
                                         uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// ---- pushes between the CTAs of a cluster: a remote shared-memory store that completes `8 bytes` on an mbarrier in the
// RECEIVER's shared memory (st.async), so the receiver waits on its own barrier and no cluster-wide barrier (whose release
// fence costs microseconds) sits on the path.
__device__ __forceinline__ uint32_t mapa_u32(uint32_t smem_addr, uint32_t cta_rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(cta_rank));
    return r;
}
__device__ __forceinline__ void st_async_b64(uint32_t dst_cluster_addr, uint64_t v, uint32_t bar_cluster_addr)
{
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];"
                 ::"r"(dst_cluster_addr), "l"(v), "r"(bar_cluster_addr)
                 : "memory");
}

__device__ __forceinline__ float warp_sum_f32(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_f64(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---- warp-level bitonic primitives on NPL x 32 uint64 keys, element e = i*32 + lane.
template <int NPL>
__device__ __forceinline__ void warp_bitonic_exchange(uint64_t (&k)[NPL], int lane, int size,
                                                      int stride)
{
    if (stride >= 32) {
        const int rs = stride >> 5;
#pragma unroll
        for (int i = 0; i < NPL; ++i) {
            if ((i & rs) == 0) {
                const int e = i * 32 + lane;
                const bool desc = (e & size) == 0;
                uint64_t a = k[i], b = k[i | rs];
                const bool swap = desc ? (a < b) : (a > b);
                if (swap) { k[i] = b; k[i | rs] = a; }
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < NPL; ++i) {
            const int e = i * 32 + lane;
            const bool desc = (e & size) == 0;
            const bool lower = (lane & stride) == 0;
            uint64_t mine = k[i];
            uint64_t other = __shfl_xor_sync(0xffffffffu, mine, stride);
            // the lower element of a descending pair keeps the max
            const bool keep_max = (desc == lower);
            k[i] = keep_max ? (mine > other ? mine : other) : (mine < other ? mine : other);
        }
    }
}

// Full sort, descending by key.
template <int NPL>
__device__ __forceinline__ void warp_bitonic_sort_desc(uint64_t (&k)[NPL], int lane)
{
    constexpr int M = NPL * 32;
    // fully unrolled so every register index is a compile-time constant (no local memory)
#pragma unroll
    for (int size = 2; size <= M; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1)
            warp_bitonic_exchange<NPL>(k, lane, size == M ? (M << 1) : size, stride);
    }
}

// k holds a bitonic sequence; sort it descending.
template <int NPL>
__device__ __forceinline__ void warp_bitonic_merge_desc(uint64_t (&k)[NPL], int lane)
{
    constexpr int M = NPL * 32;
#pragma unroll
    for (int stride = M >> 1; stride > 0; stride >>= 1)
        warp_bitonic_exchange<NPL>(k, lane, M << 1, stride);
}

// mine: sorted desc (regs); other: sorted desc list of NPL*32 keys in memory.
// Result: top NPL*32 of the union, sorted desc.
template <int NPL>
__device__ __forceinline__ void warp_merge_topk(uint64_t (&mine)[NPL], const uint64_t *other,
                                                int lane)
{
    constexpr int M = NPL * 32;
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
        const int e = i * 32 + lane;
        uint64_t o = other[M - 1 - e];
        mine[i] = mine[i] > o ? mine[i] : o;
    }
    warp_bitonic_merge_desc<NPL>(mine, lane);
}

// Per-warp running top-KC in shared memory: unsorted list + (min key, position) in registers.
template <int NPL>
struct WarpTopK {
    uint64_t *list;   // [KC]
    uint64_t tau;     // current admission threshold (0 while the list is not full)
    int count;
    int min_pos;

    __device__ __forceinline__ void init(uint64_t *l, int lane)
    {
        list = l;
        tau = 0;
        count = 0;
        min_pos = 0;
#pragma unroll
        for (int i = 0; i < NPL; ++i) list[i * 32 + lane] = CDR_EMPTY_KEY;
        __syncwarp();
    }
    __device__ __forceinline__ void refresh_min(int lane)
    {
        uint64_t m = ~0ull;
        int pos = 0;
#pragma unroll
        for (int i = 0; i < NPL; ++i) {
            uint64_t v = list[i * 32 + lane];
            if (v < m) { m = v; pos = i * 32 + lane; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            uint64_t om = __shfl_xor_sync(0xffffffffu, m, o);
            int op = __shfl_xor_sync(0xffffffffu, pos, o);
            if (om < m) { m = om; pos = op; }   // keys are unique, so no tie handling needed
        }
        tau = m;
        min_pos = pos;
    }
    // warp-uniform call
    __device__ __forceinline__ void push(uint64_t key, int lane)
    {
        constexpr int KC = NPL * 32;
        CDR_DEV_ASSERT(count >= 0 && count <= KC && min_pos >= 0 && min_pos < KC && key != CDR_EMPTY_KEY);
        if (count < KC) {
            if (lane == 0) list[count] = key;
            ++count;
            __syncwarp();
            if (count == KC) refresh_min(lane);
        } else {
            if (lane == 0) list[min_pos] = key;
            __syncwarp();
            refresh_min(lane);
        }
    }
};


// ---- peer-memory exchange of one query's list, by one CTA (the protocol is described at the top of peer.cu).
__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const void *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// Called by every thread of the CTA.  l_sc / l_id: this rank's ordered list for the query (my_n entries; shared or global
// memory, written before the call -- the function starts with a CTA barrier).  slot: the query's slot in the group's buffers.
// m_key / m_id (world * k entries), s_n (world ints) and s_cnt: shared-memory scratch.  o_*: the query's output row.
//   push  : the list goes into slot[parity][me][slot] of EVERY rank's buffer (posted NVLink writes), then flag[me][slot] =
//           epoch is published in every buffer with a system-scope release store;
//   wait  : `world` threads spin (acquire loads on LOCAL memory) until flag[r][slot] >= epoch for every r;
//   merge : the `world` lists are ranked by counting with the global ordering rule and the first k written out.
__device__ __forceinline__ void peer_exchange_cta(const PeerLink &pl, int slot, int k, const double *l_sc,
                                                  const int64_t *l_id, int my_n, uint64_t *m_key, int64_t *m_id, int *s_n,
                                                  int *s_cnt, double *o_sc, int64_t *o_id, int32_t *o_n)
{
    const uint32_t par = pl.epoch & 1u;
    const size_t slot_me = (((size_t)par * pl.world + pl.rank) * pl.max_nq + slot) * pl.entry_bytes;
    CDR_DEV_ASSERT(slot >= 0 && slot < pl.max_nq && k <= pl.max_k && slot_me + pl.entry_bytes <= pl.flags_off && my_n >= 0 &&
                   my_n <= k);
    if (threadIdx.x == 0) *s_cnt = 0;
    __syncthreads();
    for (int r = 0; r < pl.world; ++r) {
        unsigned char *dst = pl.base[r] + slot_me;
        unsigned long long *d_sc = reinterpret_cast<unsigned long long *>(dst);
        unsigned long long *d_id = d_sc + pl.max_k;
        for (int i = threadIdx.x; i < my_n; i += blockDim.x) {
            d_sc[i] = (unsigned long long)__double_as_longlong(l_sc[i]);
            d_id[i] = (unsigned long long)l_id[i];
        }
        if (threadIdx.x == 0) *reinterpret_cast<int32_t *>(d_id + pl.max_k) = my_n;
    }
    __syncthreads();
    if ((int)threadIdx.x < pl.world) {
        const int r = threadIdx.x;
        // the CTA's stores above are ordered before this release (bar.sync + cumulativity)
        __threadfence_system();
        st_release_sys(reinterpret_cast<uint32_t *>(pl.base[r] + pl.flags_off) + (size_t)pl.rank * pl.max_nq + slot, pl.epoch);
        // rank r's list for this query has landed in MY buffer
        const uint32_t *flag =
            reinterpret_cast<const uint32_t *>(pl.base[pl.rank] + pl.flags_off) + (size_t)r * pl.max_nq + slot;
        long long t0 = clock64();
        while ((int32_t)(ld_acquire_sys(flag) - pl.epoch) < 0) {
            if (clock64() - t0 > 120000000000ll) __trap();   // ~60 s: a peer never launched -> fail, do not hang
        }
        const unsigned char *src = pl.base[pl.rank] + (((size_t)par * pl.world + r) * pl.max_nq + slot) * pl.entry_bytes;
        s_n[r] = *reinterpret_cast<const volatile int32_t *>(src + (size_t)pl.max_k * 16);
    }
    __syncthreads();

    // merge: every list arrives ordered, so entry i of list r has global rank i + (entries of the other lists before it),
    // each found by a branch-free binary search (cdr_sorted_count_before); a winner's score bits are read again from the buffer
    const int tot = pl.world * k;
    int local = 0;
    for (int e = threadIdx.x; e < tot; e += blockDim.x) {
        const int r = e / k, i = e - r * k;
        const bool ok = i < s_n[r];
        const unsigned char *src = pl.base[pl.rank] + (((size_t)par * pl.world + r) * pl.max_nq + slot) * pl.entry_bytes;
        m_key[e] = ok ? cdr_order_f64_bits(ld_relaxed_sys_u64(src + (size_t)i * 8)) : 0ull;
        m_id[e] = ok ? (int64_t)ld_relaxed_sys_u64(src + (size_t)(pl.max_k + i) * 8) : INT64_MAX;
        local += ok;
    }
    if (local) atomicAdd(s_cnt, local);
    __syncthreads();
    for (int e = threadIdx.x; e < tot; e += blockDim.x) {
        const int r = e / k, i = e - r * k;
        if (i >= s_n[r]) continue;
        const uint64_t key = m_key[e];
        const int64_t id = m_id[e];
        int rank = i;
        for (int o = 0; o < pl.world; ++o)
            if (o != r) rank += cdr_sorted_count_before(m_key + o * k, m_id + o * k, s_n[o], k, key, id);
        if (rank < k) {
            // the score's bits back from its order key; the two keys that do not determine them (NaN payloads, the sign of
            // a zero) are read again from the buffer
            uint64_t bits = (key >> 63) ? (key & 0x7FFFFFFFFFFFFFFFull) : ~key;
            if (key == 0ull || key == 0x8000000000000000ull) {
                const unsigned char *src =
                    pl.base[pl.rank] + (((size_t)par * pl.world + r) * pl.max_nq + slot) * pl.entry_bytes;
                bits = ld_relaxed_sys_u64(src + (size_t)i * 8);
            }
            o_sc[rank] = __longlong_as_double((long long)bits);
            o_id[rank] = id;
        }
    }
    const int n_out = *s_cnt < k ? *s_cnt : k;
    for (int i = n_out + threadIdx.x; i < k; i += blockDim.x) {
        o_sc[i] = __longlong_as_double(0x7FF8000000000000ll);
        o_id[i] = -1;
    }
    if (threadIdx.x == 0) *o_n = n_out;
}

#endif  // __CUDACC__
