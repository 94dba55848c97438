// gemm_topk.cu -- K2: batched bf16 tensor-core scan (tcgen05 / TMEM / TMA) with a fused
// threshold-append top-k epilogue.  The score matrix S = Q . X^T never reaches HBM.
//
// Serves mode "ann" of the reference (app/retrieve.py:290-298 enables pgvector's HNSW with
// ef_search; here brute force on the 5th-gen tensor cores replaces the graph walk) for batches
// of queries.  Survivors are re-scored exactly and ordered like the exact lane (exact_scan.cu).
//
// Shapes: Q bf16 [nq_pad, D] (rows L2-normalised, zero padded to a multiple of 128),
//         X bf16 [N, D] (rows L2-normalised at ingest), both K-major.
// CTA tile: 128 queries (UMMA M) x 256 corpus rows (UMMA N), K step 64 (= one 128-byte swizzle
// atom per operand row), kStages-deep TMA->smem ring, fp32 accumulators in TMEM, two
// accumulator buffers (2 x 256 columns = all 512) so the epilogue of tile t overlaps the MMAs of
// tile t+1.
//
// Warp roles (256 or 384 threads, one CTA per SM, persistent over work items):
//   warp 0   TMA producer   : cp.async.bulk.tensor.2d (128B swizzle) A and B boxes per K block
//   warp 1   MMA issuer     : one elected lane issues tcgen05.mma.cta_group::1.kind::f16
//                             (128x256x16), tcgen05.commit releases smem stages / publishes TMEM
//   warp 2   TMEM allocator
//   warps 4-7 epilogue      : tcgen05.ld 32x32b (thread = one query row), running max of each
//   (+ 8-11 in the small,     32-column chunk against the query's threshold tau; the rare chunks
//    append-heavy segments)   that beat it stage (score,row) keys in shared memory and append them to
//                             the query's candidate list in global memory (one atomicAdd per tile).
//
// Exactness argument: every row whose bf16 score >= tau[q] is appended (complete above the
// threshold), and tau[q] is always the KC-th best score of a subset of the rows seen so far, so
// the final list contains the bf16 top-KC of the whole corpus.  The corpus is processed in
// geometrically growing segments (16, 256, 4096, ... tiles in a multiplicative permutation of
// the tile order); between segments a select kernel compacts each list to its top-KC and raises
// tau.  No row is ever multiplied twice.  List overflow (adversarial data) is flagged and the
// affected queries are re-run on the exact lane by the host wrapper.
#include "common.cuh"

#include <cuda.h>
#include <cstdlib>

int cdr_finalize_unsorted_launch(cdr_store *s, const uint64_t *lists, const uint32_t *counts, int cap,
                                 int kc, const float *q_dev, int nq, int k, bool use_bf16_rows,
                                 double *out_score, int64_t *out_id, int32_t *out_n, cudaStream_t st);

namespace {

constexpr int kBlockM = 128;
constexpr int kBlockN = 256;
constexpr int kBlockK = 64;                 // bf16 elements = 128 bytes = one swizzle atom row
constexpr int kUmmaK = 16;
constexpr int kStages = 4;
constexpr int kABytes = kBlockM * kBlockK * 2;   // 16 KB
constexpr int kBBytes = kBlockN * kBlockK * 2;   // 32 KB
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kTmemCols = 512;
// kEpi epilogue warp groups of 4 warps (one warp per TMEM lane quarter); group g reads the g-th share of a tile's 256
// accumulator columns.  One group leaves ONE warp per scheduler for the epilogue, so the slow path (a dependent chain of
// ~300 instructions per 32-column chunk with a hit) runs at the latency of a single warp: ~1 us per chunk, exposed in the
// early, append-heavy segments.  Two groups halve the columns per warp and give every scheduler two warps to interleave.
__host__ __device__ constexpr int gemm_threads(int kEpi) { return 128 + 128 * kEpi; }
// epilogue append staging: each of the 128 epilogue threads owns kBufN key slots in shared memory, laid out
// [slot][thread].  A thread's staged keys all belong to ONE query (its row of the current tile): they are flushed to that
// query's global candidate list when the tile is done -- after the accumulator buffer went back to the MMA warp -- with
// ONE atomicAdd for the whole batch (the list slots are reserved together) instead of one atomic per key.
__host__ __device__ constexpr int buf_slots(int kEpi) { return kEpi == 4 ? 8 : 16; }     // key slots per epilogue thread (<= 32 KB in all)
__host__ __device__ constexpr size_t gemm_smem(int kEpi)
{
    return (size_t)kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/ + (size_t)buf_slots(kEpi) * 128 * kEpi * 8;
}
static_assert(gemm_smem(2) <= 232448 && gemm_smem(4) <= 232448, "dynamic shared memory limit of sm_100a");

// ---- PTX wrappers (tcgen05 / TMA); forms follow the PTX ISA for sm_100a
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}
// same, multicast: the box lands at the same smem offset in every CTA of `mask`, and each
// destination CTA's mbarrier (same offset) receives the complete_tx for the bytes it got
__device__ __forceinline__ void tma_load_2d_mc(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar,
                                               uint16_t mask)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster "
        "[%0], [%1, {%2, %3}], [%4], %5;"
        ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "h"(mask)
        : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap *map)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t cols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// ---- cta_group::2 ("2-SM MMA") forms: one MMA spans the CTA pair (M = 256), each CTA holds its
// own 128 A rows and HALF of the B tile; the leader CTA (rank 0) issues, both CTAs' tensor cores work.
__device__ __forceinline__ uint32_t mapa_u32(uint32_t smem_addr, uint32_t cta_rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(cta_rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr)
{
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load into OWN shared memory whose completion bytes are credited to the LEADER CTA's mbarrier
__device__ __forceinline__ void tma_load_2d_2sm(void *dst, const CUtensorMap *map, int c0, int c1, uint32_t leader_bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(leader_bar)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t *dst_smem, uint32_t cols)
{
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t cols)
{
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_commit_2sm_mc(uint64_t *bar, uint16_t mask)
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask)
                 : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on the barrier at the same offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_mc(uint64_t *bar, uint16_t mask)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask)
                 : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 x bf16 -> fp32, M128 N256 K16
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major, 128B-swizzled operand tile: 8-row groups are 1024 B apart (SBO), LBO unused.
__device__ __forceinline__ uint64_t make_smem_desc(const void *tile)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_u32(tile) & 0x3FFFFu) >> 4);      // start address, 16-byte units
    d |= (uint64_t)0 << 16;                                  // leading byte offset (unused)
    d |= (uint64_t)(1024 >> 4) << 32;                        // stride byte offset
    d |= (uint64_t)1 << 46;                                  // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                                  // SWIZZLE_128B
    return d;
}
constexpr uint32_t kInstrDesc = (1u << 4)                    // D format: fp32
                                | (1u << 7)                  // A format: bf16
                                | (1u << 10)                 // B format: bf16
                                | ((uint32_t)(kBlockN >> 3) << 17)
                                | ((uint32_t)(kBlockM >> 4) << 24);   // A, B K-major (bits 15,16 = 0)
constexpr uint32_t kInstrDesc2Sm = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kBlockN >> 3) << 17)
                                   | ((uint32_t)((2 * kBlockM) >> 4) << 24);   // M = 256 across the CTA pair

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32])
{
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// 3-input maximum (SASS FMNMX3 on sm_100a): halves the max chain of the epilogue's fast path.  NaN operands are
// ignored like fmaxf's (the result is NaN only when all three are).
__device__ __forceinline__ float fmax3(float a, float b, float c)
{
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

struct GemmParams {
    const float *tau;          // [nq_pad] admission threshold per query as a score (-inf => take everything): fast test
    const uint64_t *tau_key;   // [nq_pad] the same threshold as a packed key (0 => take everything): exact test
    uint64_t *lists;           // [nq, cap] candidate keys
    uint32_t *counts;          // [nq]
    const uint32_t *allow;     // nullable row bitmap
    int64_t n_rows;
    int nq;                    // real queries
    int m_tiles;               // nq_pad / 128
    int cap;
    int64_t n_tiles_total;     // ceil(n_rows / 256)
    int64_t perm_mul;          // tile permutation multiplier (coprime with n_tiles_total)
    int64_t tile_begin, tile_end;   // permuted tile index range of this segment
    int k_blocks;              // D / 64
    int tile_major;            // 1: a cluster walks all query tiles of one corpus tile back to back (see item_of)
    unsigned epi_sleep_ns;     // epilogue warps sleep this long between polls of the accumulator barrier (0 = spin)
    int debug_no_append;       // measurement aid (CADENCE_K2_DRYRUN=1): treat tau as +inf => pure GEMM + max
};

// A corpus tile (kBlockN = 256 rows = 8 bitmap words) none of whose rows passes the filter is skipped by all
// three roles (same deterministic test, so their pipeline counters stay in step): no TMA, no MMA, no epilogue.
// Filters with block structure -- a date range over time-ordered rows, a set of calls -- leave most tiles empty.
__device__ __forceinline__ bool tile_has_allowed_rows(const uint32_t *allow, int64_t nt, int64_t n_rows)
{
    if (allow == nullptr) return true;
    const int64_t w0 = nt * (kBlockN / 32);
    const int64_t words = (n_rows + 31) >> 5;
    uint32_t any = 0;
#pragma unroll
    for (int i = 0; i < kBlockN / 32; ++i)
        if (w0 + i < words) any |= __ldg(&allow[w0 + i]);
    return any != 0;
}

// Flush one epilogue thread's staged keys to its query's global candidate list: one atomic reserves the slots, then the
// stores.  Deliberately NOT inlined: the epilogue's hot loop must stay small enough for the instruction cache.
template <int kBufN, int kStride>
__device__ __noinline__ void flush_staged(uint32_t *counts, uint64_t *lists, int cap, const uint64_t *buf_key, int es,
                                          int nbuf, int q)
{
    CDR_DEV_ASSERT(nbuf > 0 && nbuf <= kBufN && es >= 0 && es < kStride && q >= 0);
    const uint32_t base = atomicAdd(&counts[q], (uint32_t)nbuf);     // counts may pass cap: that IS the overflow signal
    uint64_t *dst = lists + (size_t)q * cap;
#pragma unroll
    for (int i = 0; i < kBufN; ++i)
        if (i < nbuf && base + (uint32_t)i < (uint32_t)cap) dst[base + i] = buf_key[i * kStride + es];
}

// kCluster == 2: the two CTAs of a cluster work on the same corpus tile with different query
// tiles; each loads half of the B (corpus) box and multicasts it to both, halving the L2->SM
// traffic of the large operand.  A stage may be refilled only when BOTH CTAs' MMAs have drained
// it, so the MMA warps commit to the empty barrier of both CTAs.
// k2Sm (requires kCluster == 2): cta_group::2 MMAs.  Each CTA stages its own A tile and only its half
// of the B tile (32 KB per stage instead of 48 KB => 6 stages instead of 4, and a third less shared
// memory read per MMA), TMA completions of both CTAs are credited to the leader's "full" barrier,
// the leader's MMA warp issues for the pair and commits to both CTAs' barriers, and both CTAs'
// epilogue warps hand their TMEM buffers back to the leader.
template <int kCluster, bool k2Sm, int kEpi>
__global__ void __launch_bounds__(gemm_threads(kEpi), 1)
gemm_topk_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_x,
                 const GemmParams p)
{
    extern __shared__ unsigned char smem_raw[];
    // 128B swizzle needs 1024-byte aligned tiles
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    static_assert(!k2Sm || kCluster == 2, "2-SM MMA needs a CTA pair");
    static_assert(kEpi == 1 || kEpi == 2 || kEpi == 4, "epilogue warp groups");
    constexpr int kBufN = buf_slots(kEpi);
    constexpr int kEpiThreads = 128 * kEpi;                               // staging stride: [slot][epilogue thread]
    constexpr int kChunksPer = (kBlockN / 32) / kEpi;                     // 32-column chunks per epilogue warp and tile
    constexpr int kStg = k2Sm ? 6 : kStages;                              // ring depth
    constexpr int kBLocal = k2Sm ? kBBytes / 2 : kBBytes;                  // B bytes staged per CTA per stage
    constexpr int kStgBytes = kABytes + kBLocal;
    static_assert(kStg * kStgBytes == kStages * kStageBytes, "both modes use the same tile memory");
    unsigned char *tiles = smem;
    uint64_t *bars = reinterpret_cast<uint64_t *>(tiles + (size_t)kStages * kStageBytes);
    uint64_t *full_bar = bars;                    // [kStg]
    uint64_t *empty_bar = bars + kStg;            // [kStg]
    uint64_t *tmem_full = bars + 2 * kStg;        // [2]
    uint64_t *tmem_empty = bars + 2 * kStg + 2;   // [2]
    uint32_t *tmem_base_slot = reinterpret_cast<uint32_t *>(bars + 2 * kStg + 4);
    uint64_t *buf_key = reinterpret_cast<uint64_t *>(reinterpret_cast<unsigned char *>(bars) + 256);   // [kBufN][kEpiThreads]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&map_q);
        prefetch_tmap(&map_x);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStg; ++s) {
            mbar_init(&full_bar[s], k2Sm ? 2 : 1);               // 2-SM: leader's expect_tx arrive + peer's arrive
            mbar_init(&empty_bar[s], k2Sm ? 1 : kCluster);       // 2-SM: one multicast commit from the leader
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tmem_full[b], 1);
            mbar_init(&tmem_empty[b], (k2Sm ? 8 : 4) * kEpi);    // one arrival per epilogue warp (of both CTAs)
        }
        fence_mbar_init();
    }
    if (warp == 2) {
        if (k2Sm) tmem_alloc_2sm(tmem_base_slot, kTmemCols);
        else tmem_alloc(tmem_base_slot, kTmemCols);
    }
    tc_fence_before();
    __syncthreads();
    if (kCluster > 1) cluster_sync_all();     // peer barriers are initialised before any remote arrive
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_slot;

    const uint32_t crank = kCluster > 1 ? cluster_ctarank() : 0u;
    const int mt_per = p.m_tiles / kCluster;             // query tiles per CTA of the cluster
    const int64_t seg_tiles = p.tile_end - p.tile_begin;
    const int64_t G = gridDim.x / kCluster;
    const int64_t w0 = blockIdx.x / kCluster;
    // Work items of this cluster, i = 0 .. n_my-1 (both CTAs of a cluster see the same sequence):
    //   item-major (tile_major = 0): global item w = w0 + i*G -> (corpus tile w / mt_per, query tile w % mt_per);
    //       the mt_per clusters that share a corpus tile run concurrently (best balance, small segments)
    //   tile-major (tile_major = 1): corpus tile w0 + (i / mt_per)*G, query tile i % mt_per; the cluster reads a
    //       corpus tile from DRAM once and re-reads it from L2 for the remaining query tiles
    int64_t n_my;
    if (p.tile_major) n_my = (seg_tiles > w0 ? (seg_tiles - w0 + G - 1) / G : 0) * mt_per;
    else n_my = (seg_tiles * mt_per > w0) ? (seg_tiles * mt_per - w0 + G - 1) / G : 0;
    auto item_of = [&](int64_t i, int64_t &tj, int &mt) {
        if (p.tile_major) {
            tj = p.tile_begin + w0 + (i / mt_per) * G;
            mt = (int)(i % mt_per) * kCluster + (int)crank;
        } else {
            const int64_t w = w0 + i * G;
            tj = p.tile_begin + w / mt_per;
            mt = (int)(w % mt_per) * kCluster + (int)crank;
        }
    };

    if (warp == 0) {
        // ---------------------------------------------------------------- TMA producer
        if (lane == 0) {
            uint32_t it = 0;
            for (int64_t i = 0; i < n_my; ++i) {
                int64_t tj; int mt;
                item_of(i, tj, mt);
                const int64_t nt = (tj * p.perm_mul) % p.n_tiles_total;
                if (!tile_has_allowed_rows(p.allow, nt, p.n_rows)) continue;
                for (int kb = 0; kb < p.k_blocks; ++kb, ++it) {
                    const int s = it % kStg;
                    const uint32_t ph = (it / kStg) & 1u;
                    mbar_wait(&empty_bar[s], ph ^ 1u);
                    unsigned char *a = tiles + (size_t)s * kStgBytes;
                    if (k2Sm) {
                        const uint32_t leader_bar = mapa_u32(smem_u32(&full_bar[s]), 0u);
                        if (crank == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * kStgBytes);   // both CTAs' bytes
                        else mbar_arrive_cluster(leader_bar);
                        tma_load_2d_2sm(a, &map_q, kb * kBlockK, mt * kBlockM, leader_bar);
                        tma_load_2d_2sm(a + kABytes, &map_x, kb * kBlockK,
                                        (int)(nt * kBlockN) + (int)crank * (kBlockN / 2), leader_bar);
                        continue;
                    }
                    if (p.debug_no_append == 7) {                 // timing aid: no operand traffic at all (MMA issue alone)
                        mbar_arrive(&full_bar[s]);
                        continue;
                    }
                    if (p.debug_no_append == 6) {                 // timing aid: corpus tiles only, the query tile is not re-read
                        mbar_arrive_expect_tx(&full_bar[s], kBBytes);
                    } else {
                        mbar_arrive_expect_tx(&full_bar[s], kStageBytes);
                        tma_load_2d(a, &map_q, kb * kBlockK, mt * kBlockM, &full_bar[s]);
                    }
                    if (kCluster == 1) {
                        tma_load_2d(a + kABytes, &map_x, kb * kBlockK, (int)(nt * kBlockN), &full_bar[s]);
                    } else {
                        // this CTA's half of the corpus box (128 rows), delivered to both CTAs
                        const int half = kBlockN / kCluster;
                        tma_load_2d_mc(a + kABytes + (size_t)crank * (kBBytes / kCluster), &map_x, kb * kBlockK,
                                       (int)(nt * kBlockN) + (int)crank * half, &full_bar[s], (uint16_t)((1u << kCluster) - 1u));
                    }
                }
            }
        }
    } else if (warp == 1 && !(k2Sm && crank != 0)) {
        // ---------------------------------------------------------------- MMA issuer (2-SM: leader CTA only)
        uint32_t it = 0, tile_no = 0;
        for (int64_t i = 0; i < n_my; ++i) {
            {
                int64_t tj; int mt;
                item_of(i, tj, mt);
                if (!tile_has_allowed_rows(p.allow, (tj * p.perm_mul) % p.n_tiles_total, p.n_rows)) continue;
            }
            const uint32_t buf = tile_no & 1u;
            const uint32_t use = tile_no >> 1;                    // how many times this buffer was used before
            mbar_wait(&tmem_empty[buf], (use & 1u) ^ 1u);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + buf * kBlockN;
            for (int kb = 0; kb < p.k_blocks; ++kb, ++it) {
                const int s = it % kStg;
                const uint32_t ph = (it / kStg) & 1u;
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                if (lane == 0) {
                    const unsigned char *a = tiles + (size_t)s * kStgBytes;
                    const uint64_t adesc = make_smem_desc(a);
                    const uint64_t bdesc = make_smem_desc(a + kABytes);
#pragma unroll
                    for (int k4 = 0; k4 < kBlockK / kUmmaK; ++k4) {
                        // advance 16 bf16 = 32 bytes along K inside the swizzle atom: +2 in 16-byte units
                        if (k2Sm)
                            umma_bf16_2sm(tmem_d, adesc + (uint64_t)(k4 * 2), bdesc + (uint64_t)(k4 * 2), kInstrDesc2Sm,
                                          (kb | k4) != 0 ? 1u : 0u);
                        else
                            umma_bf16(tmem_d, adesc + (uint64_t)(k4 * 2), bdesc + (uint64_t)(k4 * 2), kInstrDesc,
                                      (kb | k4) != 0 ? 1u : 0u);
                    }
                    // smem stage free when these MMAs retire (in every CTA that TMA-writes it)
                    if (k2Sm) umma_commit_2sm_mc(&empty_bar[s], (uint16_t)0x3);
                    else if (kCluster == 1) umma_commit(&empty_bar[s]);
                    else umma_commit_mc(&empty_bar[s], (uint16_t)((1u << kCluster) - 1u));
                    if (kb == p.k_blocks - 1) {
                        if (k2Sm) umma_commit_2sm_mc(&tmem_full[buf], (uint16_t)0x3);
                        else umma_commit(&tmem_full[buf]);
                    }
                }
                __syncwarp();
            }
            ++tile_no;
        }
    } else if (warp >= 4) {
        // ---------------------------------------------------------------- epilogue
        const int ew = warp & 3;                                  // TMEM lane quarter of this warp
        const int et = ew * 32 + lane;                            // query row of the tile, 0..127
        const int eg = (warp - 4) >> 2;                           // column group of this warp, 0..kEpi-1
        const int es = eg * 128 + et;                             // staging column of this thread
        int nbuf = 0;
        uint32_t tile_no = 0;
        for (int64_t i = 0; i < n_my; ++i) {
            int64_t tj; int mt;
            item_of(i, tj, mt);
            const int64_t nt = (tj * p.perm_mul) % p.n_tiles_total;
            if (!tile_has_allowed_rows(p.allow, nt, p.n_rows)) continue;
            const int64_t row0 = nt * kBlockN;
            const int q = mt * kBlockM + et;
            const bool q_ok = q < p.nq;
            const bool live = q_ok && (p.debug_no_append == 0 || p.debug_no_append == 2 || p.debug_no_append == 3);
            const float tau = live ? __ldg(&p.tau[q]) : __int_as_float(0x7f800000);   // +inf: never passes
            const uint64_t tau_key = live ? __ldg(&p.tau_key[q]) : ~0ull;
            int64_t lim = p.n_rows - row0;                        // valid columns in this tile
            if (lim > kBlockN) lim = kBlockN;
            const uint32_t buf = tile_no & 1u;
            const uint32_t use = tile_no >> 1;
            CDR_DEV_ASSERT(nt >= 0 && nt < p.n_tiles_total && mt >= 0 && mt < p.m_tiles && row0 < p.n_rows);
            mbar_wait_backoff(&tmem_full[buf], use & 1u, p.epi_sleep_ns);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + buf * kBlockN;
#pragma unroll 1
            for (int c = eg * kChunksPer; c < (eg + 1) * kChunksPer; ++c) {
                __syncwarp();
                if (p.debug_no_append >= 5) continue;             // timing aid: no epilogue work at all (mainloop alone)
                float v[32];
                tmem_ld32(taddr + c * 32, v);
                if (p.debug_no_append == 4) {                     // timing aid: TMEM reads only, no maxima
                    if (v[0] == 12345.678f && v[31] == 8765.4321f) buf_key[es] = 1;   // keeps the load alive
                    continue;
                }
                if (c * 32 >= lim) continue;                      // warp-uniform (tail tile)
                // fast path: the maximum of the chunk as four 8-column sub-maxima (3-input max: 18 instead of 31
                // operations per chunk) against the query's threshold.  !(m < tau) also holds for a NaN maximum
                // (a chunk of NaN scores: rows with a zero or non-finite norm), which the slow path sorts out.
                float ms[4];
#pragma unroll
                for (int s4 = 0; s4 < 4; ++s4) {
                    const float *w = v + 8 * s4;
                    ms[s4] = fmaxf(fmax3(fmax3(w[0], w[1], w[2]), fmax3(w[3], w[4], w[5]), w[6]), w[7]);
                }
                const float m = fmaxf(fmax3(ms[0], ms[1], ms[2]), ms[3]);
                if (!(m < tau)) {
                    // slow path (rare): only the sub-chunks whose maximum passed are looked at element by element;
                    // admission is decided on the packed key (score desc, row asc), like the exact lane: NaN scores
                    // stay eligible below every real score, ties with the threshold are cut by row order
                    uint32_t allow_w = 0xFFFFFFFFu;
                    // 32 consecutive rows starting at a multiple of 32: exactly one bitmap word
                    if (p.allow != nullptr) allow_w = __ldg(&p.allow[(row0 + c * 32) >> 5]);
                    const int rem = (int)(lim - c * 32);
                    if (rem < 32) allow_w &= (1u << rem) - 1u;
                    if (p.debug_no_append == 2) allow_w = 0;       // timing aid: tests run, nothing staged
#pragma unroll
                    for (int s4 = 0; s4 < 4; ++s4) {
                        if (!(ms[s4] < tau)) {
                            if (p.debug_no_append == 3) nbuf = 0;  // timing aid: stage, never flush
                            if (nbuf > kBufN - 8) {                // make room for up to 8 hits
                                flush_staged<kBufN, kEpiThreads>(p.counts, p.lists, p.cap, buf_key, es, nbuf, q);
                                nbuf = 0;
                            }
                            const uint32_t row_lo = 0xFFFFFFFFu - (uint32_t)(row0 + c * 32 + 8 * s4);   // key low word of column 8*s4
#pragma unroll
                            for (int j = 8 * s4; j < 8 * s4 + 8; ++j) {
                                if (!(v[j] < tau) && ((allow_w >> j) & 1u)) {
                                    // above the threshold: in.  Equal to it (or NaN): the packed key decides (row order / NaN code)
                                    const uint64_t key = ((uint64_t)cdr_order_f32(v[j]) << 32) | (uint64_t)(row_lo - (uint32_t)(j - 8 * s4));
                                    if (v[j] > tau || key > tau_key) {
                                        CDR_DEV_ASSERT(nbuf >= 0 && nbuf < kBufN && q < p.nq && row0 + c * 32 + j < p.n_rows);
                                        buf_key[nbuf * kEpiThreads + es] = key;
                                        ++nbuf;
                                    }
                                }
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (k2Sm && crank != 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty[buf]), 0u));   // leader's barrier
                else mbar_arrive(&tmem_empty[buf]);
            }
            // The tile's keys go out NOW, after the accumulator buffer went back to the MMA warp (the atomic's round trip
            // overlaps the next tile's MMAs instead of holding TMEM): the next tile is another query's.
            if (nbuf > 0 && p.debug_no_append != 3) flush_staged<kBufN, kEpiThreads>(p.counts, p.lists, p.cap, buf_key, es, nbuf, q);
            nbuf = 0;
            ++tile_no;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (kCluster > 1) cluster_sync_all();     // no CTA exits while its peer may still multicast into it
    if (warp == 2) {
        tc_fence_after();
        if (k2Sm) tmem_dealloc_2sm(tmem_base, kTmemCols);
        else tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ---- query preparation: q_bf16[q,:] = RN(q / ||q||), zero rows for padding; tau = -inf
__global__ void prep_queries_kernel(const float *q, __nv_bfloat16 *out, float *tau, uint64_t *tau_key, uint32_t *counts,
                                    uint32_t *overflow, int nq, int nq_pad, int dim)
{
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= nq_pad) return;
    __nv_bfloat16 *o = out + (size_t)row * dim;
    if (row >= nq) {
        for (int i = lane; i < dim; i += 32) o[i] = __float2bfloat16_rn(0.f);
        if (lane == 0) {
            tau[row] = __int_as_float(0x7f800000);
            tau_key[row] = ~0ull;
        }
        return;
    }
    const float *r = q + (size_t)row * dim;
    float n2 = 0.f;
    for (int i = lane; i < dim; i += 32) n2 = fmaf(r[i], r[i], n2);
    n2 = warp_sum_f32(n2);
    const float inv = n2 > 0.f ? __fdiv_rn(1.0f, __fsqrt_rn(n2)) : 0.f;
    for (int i = lane; i < dim; i += 32) o[i] = __float2bfloat16_rn(r[i] * inv);
    if (lane == 0) {
        tau[row] = __int_as_float(0xff800000);   // -inf
        tau_key[row] = 0ull;                     // every key, NaN scores included, is admitted until the list is full
        counts[row] = 0;
        overflow[row] = 0;
    }
}

__device__ __forceinline__ float key_score(uint64_t key)
{
    const uint32_t o = (uint32_t)(key >> 32);
    return __uint_as_float((o & 0x80000000u) ? (o ^ 0x80000000u) : ~o);
}

// ---- between segments: compact list[q] to its top-KC (sorted desc), raise tau[q]
template <int NPL>
__global__ void __launch_bounds__(256) select_compact_kernel(uint64_t *lists, uint32_t *counts, float *tau,
                                                             uint64_t *tau_key, uint32_t *overflow, int cap)
{
    constexpr int KC = NPL * 32;
    __shared__ uint64_t s_lists[8 * KC];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = blockIdx.x;
    uint64_t *list = lists + (size_t)q * cap;
    uint32_t n = counts[q];
    if (n > (uint32_t)cap) {
        if (threadIdx.x == 0) overflow[q] = 1;
        n = cap;
    }
    uint64_t k[NPL];
#pragma unroll
    for (int i = 0; i < NPL; ++i) k[i] = CDR_EMPTY_KEY;
    const int chunks = (int)((n + KC - 1) / KC);
    for (int ch = warp; ch < chunks; ch += 8) {
        uint64_t c[NPL];
#pragma unroll
        for (int i = 0; i < NPL; ++i) {
            const uint32_t e = (uint32_t)ch * KC + i * 32 + lane;
            c[i] = e < n ? list[e] : CDR_EMPTY_KEY;
        }
        warp_bitonic_sort_desc<NPL>(c, lane);
        // merge: k = top KC of (k U c); c reversed is read through shared memory
#pragma unroll
        for (int i = 0; i < NPL; ++i) s_lists[warp * KC + i * 32 + lane] = c[i];
        __syncwarp();
        warp_merge_topk<NPL>(k, s_lists + warp * KC, lane);
        __syncwarp();
    }
#pragma unroll
    for (int i = 0; i < NPL; ++i) s_lists[warp * KC + i * 32 + lane] = k[i];
#pragma unroll
    for (int step = 1; step < 8; step <<= 1) {
        __syncthreads();
        if ((warp & (2 * step - 1)) == 0) warp_merge_topk<NPL>(k, s_lists + (warp + step) * KC, lane);
        __syncthreads();
        if ((warp & (2 * step - 1)) == 0) {
#pragma unroll
            for (int i = 0; i < NPL; ++i) s_lists[warp * KC + i * 32 + lane] = k[i];
        }
    }
    __syncthreads();   // every read of list[] happened before the merges above
    if (warp == 0) {
#pragma unroll
        for (int i = 0; i < NPL; ++i) list[i * 32 + lane] = k[i];
        const uint64_t last = __shfl_sync(0xffffffffu, k[NPL - 1], 31);   // element KC-1
        if (lane == 0) {
            counts[q] = n < (uint32_t)KC ? n : (uint32_t)KC;
            if (last != CDR_EMPTY_KEY) {
                // the list is full: its KC-th key is the new threshold.  While that key is a NaN score (fewer than KC
                // real scores seen so far) the float test must let every real score through.
                tau_key[q] = last;
                tau[q] = (uint32_t)(last >> 32) > 1u ? key_score(last) : __int_as_float(0xff800000);
            }
        }
    }
}

// ---- after the last segment: the queries whose candidate list overflowed in any segment (lost candidates), as a
// compact list for the conditional exact-lane re-run; cnt[r] = live slots of round r (kRedoSlots slots per round)
constexpr int kRedoSlots = 1024;
__global__ void __launch_bounds__(1024) redo_list_kernel(const uint32_t *overflow, const uint32_t *counts, int cap, int nq,
                                                         int *idx, int *cnt, int slots, int rounds)
{
    __shared__ int s_n;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    for (int q = threadIdx.x; q < nq; q += blockDim.x)
        if (overflow[q] != 0u || counts[q] > (uint32_t)cap) idx[atomicAdd(&s_n, 1)] = q;
    __syncthreads();
    for (int r = threadIdx.x; r < rounds; r += blockDim.x) {
        const int left = s_n - r * slots;
        cnt[r] = left < 0 ? 0 : (left > slots ? slots : left);
    }
}

typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int get_encode_fn(encode_tiled_fn *out)
{
    static encode_tiled_fn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CDR_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
        CDR_REQUIRE(p != nullptr && qres == cudaDriverEntryPointSuccess, CDR_ERR_CUDA,
                    "cuTensorMapEncodeTiled not available from the driver");
        fn = (encode_tiled_fn)p;
    }
    *out = fn;
    return CDR_OK;
}

int make_map(CUtensorMap *map, const void *base, int64_t rows, int dim, int box_rows)
{
    encode_tiled_fn enc;
    int rc = get_encode_fn(&enc);
    if (rc != CDR_OK) return rc;
    cuuint64_t gdim[2] = {(cuuint64_t)dim, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)dim * 2};
    cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CDR_REQUIRE(r == CUDA_SUCCESS, CDR_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%lld dim=%d", (int)r,
                (long long)rows, dim);
    return CDR_OK;
}

int64_t gcd64(int64_t a, int64_t b) { while (b) { int64_t t = a % b; a = b; b = t; } return a; }

// CADENCE_K2_CLUSTER selects the GEMM variant (A/B measurements): 1 = independent CTAs,
// 2 = multicast the corpus tile across a 2-CTA cluster, 3 = cta_group::2 MMAs (2-SM; default since round 2: three
// interleaved runs each on one box, 10M x 1024 queries: 16.73-16.87 ms of GEMM per batch vs 16.94-17.11 for the multicast
// form, profiles/r02/k2_epilogue_groups/), 4 = the corpus tile multicast across a 4-CTA cluster (4 query tiles).
int g_k2_cluster = [] {
    const char *e = getenv("CADENCE_K2_CLUSTER");
    const int v = e ? atoi(e) : 3;
    return (v >= 1 && v <= 4) ? v : 3;
}();

}  // namespace

extern "C" int32_t cdr_search_batch_bf16(cdr_store *s, const float *q_dev, int32_t nq, int32_t k,
                                         const uint32_t *allow_dev, double *out_score_dev,
                                         int64_t *out_id_dev, int32_t *out_n_dev, void *stream)
{
    CDR_REQUIRE(s != nullptr, CDR_ERR_INVALID, "cdr_search_batch_bf16: store is NULL");
    CDR_REQUIRE(s->finalized, CDR_ERR_STATE, "cdr_search_batch_bf16: store not finalized");
    CDR_REQUIRE(s->emb_bf16 != nullptr, CDR_ERR_STATE,
                "cdr_search_batch_bf16: store has no bf16 rows (created without CDR_STORE_BF16)");
    CDR_REQUIRE(nq >= 0 && (nq == 0 || q_dev), CDR_ERR_INVALID, "cdr_search_batch_bf16: bad query batch");
    CDR_REQUIRE(k >= 1 && k <= 192, CDR_ERR_UNSUPPORTED, "cdr_search_batch_bf16: k=%d outside [1,192]", k);
    CDR_REQUIRE(out_score_dev && out_id_dev && out_n_dev, CDR_ERR_INVALID, "cdr_search_batch_bf16: outputs required");
    CDR_REQUIRE(s->dim % kBlockK == 0, CDR_ERR_UNSUPPORTED, "cdr_search_batch_bf16: dim %% 64 != 0");
    if (nq == 0) return CDR_OK;
    // the epilogue stages 16-bit query indices: larger batches run as chunks of 16384 queries
    constexpr int kMaxBatch = 16384;
    if (nq > kMaxBatch) {
        for (int q0 = 0; q0 < nq; q0 += kMaxBatch) {
            const int m = nq - q0 < kMaxBatch ? nq - q0 : kMaxBatch;
            int rc = cdr_search_batch_bf16(s, q_dev + (size_t)q0 * s->dim, m, k, allow_dev,
                                           out_score_dev + (size_t)q0 * k, out_id_dev + (size_t)q0 * k,
                                           out_n_dev + q0, stream);
            if (rc != CDR_OK) return rc;
        }
        return CDR_OK;
    }
    DeviceGuard g(s->device);
    cudaStream_t st = (cudaStream_t)stream;
    std::unique_lock<std::mutex> lk(s->mu);
    return cdr_batch_bf16_launch(s, s->ws[st], q_dev, nq, k, allow_dev, out_score_dev, out_id_dev, out_n_dev, st);
}

// The lane itself; the caller holds the store mutex and has validated the arguments (nq <= 16384, k <= 192,
// bf16 rows resident).  Also the dense lane of the fused hybrid call for mode "ann" (csrc/hybrid.cu).
int cdr_batch_bf16_launch(cdr_store *s, ScanWorkspace &ws, const float *q_dev, int nq, int k, const uint32_t *allow_dev,
                          double *out_score_dev, int64_t *out_id_dev, int32_t *out_n_dev, cudaStream_t st)
{
    const int kc = k <= 64 ? 128 : 256;          // candidates kept per query (>= k + 64)
    int cap = 32 * kc;
    // test hook: a smaller list capacity makes the first segment overflow, so the tests can drive the device-side
    // re-run of overflowed queries (read per call; costs nothing next to a batch)
    if (const char *e = getenv("CADENCE_K2_TEST_CAP")) {
        const int v = atoi(e);
        if (v >= kc && v < cap) cap = v;
    }
    const int nq_pad = (nq + kBlockM - 1) / kBlockM * kBlockM;
    const int dim = s->dim;
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t b_q = up((size_t)nq_pad * dim * 2), b_tau = up((size_t)nq_pad * 4), b_cnt = up((size_t)nq_pad * 4);
    const size_t b_ovf = up((size_t)nq_pad * 4), b_lists = up((size_t)nq * cap * 8), b_tk = up((size_t)nq_pad * 8);
    const size_t b_redo = up(((size_t)nq + 64) * 4);
    if (cdr_ws_reserve(&ws.gemm_ws, &ws.gemm_ws_bytes, b_q + b_tau + b_tk + b_cnt + b_ovf + b_redo + b_lists) != CDR_OK) return CDR_ERR_OOM;
    unsigned char *c = (unsigned char *)ws.gemm_ws;
    __nv_bfloat16 *q_bf16 = (__nv_bfloat16 *)c; c += b_q;
    float *tau = (float *)c; c += b_tau;
    uint64_t *tau_key = (uint64_t *)c; c += b_tk;
    uint32_t *counts = (uint32_t *)c; c += b_cnt;
    uint32_t *overflow = (uint32_t *)c; c += b_ovf;
    int *redo_cnt = (int *)c;                    // [<= 64] live slots per re-run round
    int *redo_idx = redo_cnt + 64;               // [nq] overflowed queries
    c += b_redo;
    uint64_t *lists = (uint64_t *)c;

    prep_queries_kernel<<<(nq_pad + 7) / 8, 256, 0, st>>>(q_dev, q_bf16, tau, tau_key, counts, overflow, nq, nq_pad, dim);
    CDR_LAUNCH_CHECK();

    CUtensorMap map_q, map_x;
    int rc = make_map(&map_q, q_bf16, nq_pad, dim, kBlockM);
    if (rc != CDR_OK) return rc;
    // B-tile multicast across a 2-CTA cluster needs an even number of query tiles
    const int m_tiles = nq_pad / kBlockM;
    int cluster = (m_tiles % 2 == 0 && g_k2_cluster >= 2) ? 2 : 1;
    if (g_k2_cluster == 4 && m_tiles % 4 == 0) cluster = 4;
    const bool two_sm = cluster == 2 && g_k2_cluster == 3;
    rc = make_map(&map_x, s->emb_bf16, s->n_rows, dim, kBlockN / cluster);
    if (rc != CDR_OK) return rc;

    // epilogue warp groups (see gemm_threads()): two in the small, append-heavy segments, one in the large ones, where the
    // appends are rare and the extra warps only add barrier polling beside the MMA issuer (measured: segments 2-3 of a
    // 10M x 1024 batch 0.15 / 0.30 -> 0.08 / 0.19 ms with two groups, the 7.5 M-row segment +1 %).  CADENCE_K2_EPI = 1 | 2 | 4
    // forces one form for every segment (A/B aid).
    static const int epi_forced = [] {
        const char *e = getenv("CADENCE_K2_EPI");
        const int v = e ? atoi(e) : 0;
        return (v == 1 || v == 2 || v == 4) ? v : 0;
    }();
    typedef void (*gemm_fn)(const CUtensorMap, const CUtensorMap, const GemmParams);
    const int form = two_sm ? 2 : (cluster == 4 ? 3 : (cluster == 2 ? 1 : 0));     // c1, c2 multicast, c2 2-SM MMA, c4 multicast
    static const gemm_fn fns[4][3] = {
        {gemm_topk_kernel<1, false, 1>, gemm_topk_kernel<1, false, 2>, gemm_topk_kernel<1, false, 4>},
        {gemm_topk_kernel<2, false, 1>, gemm_topk_kernel<2, false, 2>, gemm_topk_kernel<2, false, 4>},
        {gemm_topk_kernel<2, true, 1>, gemm_topk_kernel<2, true, 2>, gemm_topk_kernel<2, true, 4>},
        {gemm_topk_kernel<4, false, 1>, gemm_topk_kernel<4, false, 2>, gemm_topk_kernel<4, false, 4>}};
    static std::mutex attr_mu;
    static bool attr_done[64] = {false};
    std::unique_lock<std::mutex> attr_lock(attr_mu);
    if (!attr_done[s->device & 63]) {
        for (int f = 0; f < 4; ++f)
            for (int e = 0; e < 3; ++e)
                CDR_CUDA(cudaFuncSetAttribute((const void *)fns[f][e], cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)gemm_smem(1 << e)));
        attr_done[s->device & 63] = true;
    }
    attr_lock.unlock();

    GemmParams p;
    p.tau = tau;
    p.tau_key = tau_key;
    p.lists = lists;
    p.counts = counts;
    p.allow = allow_dev ? allow_dev : (s->any_invalid ? s->valid : nullptr);
    p.n_rows = s->n_rows;
    p.nq = nq;
    p.m_tiles = m_tiles;
    p.cap = cap;
    p.n_tiles_total = (s->n_rows + kBlockN - 1) / kBlockN;
    p.k_blocks = dim / kBlockK;
    static const int dryrun = [] { const char *e = getenv("CADENCE_K2_DRYRUN"); return e ? atoi(e) : 0; }();
    p.debug_no_append = dryrun;
    static const unsigned epi_sleep = [] { const char *e = getenv("CADENCE_K2_EPI_SLEEP"); return e ? (unsigned)atoi(e) : 0u; }();
    p.epi_sleep_ns = epi_sleep;
    // CADENCE_K2_ORDER: 0 = item-major everywhere, 1 (default) = tile-major on large segments.  Measured at
    // 10M rows x 1024 queries: 16.43 vs 16.58-16.64 ms per batch, DRAM traffic 1.00x algorithmic either way
    // (profiles/r01/k2_dram_per_launch_order{0,1}.csv).
    static const int k2_order = [] { const char *e = getenv("CADENCE_K2_ORDER"); return e ? atoi(e) : 1; }();
    // multiplicative permutation of the tile order so every segment samples the whole corpus
    int64_t mul = 1;
    if (p.n_tiles_total > 2) {
        mul = (int64_t)((double)p.n_tiles_total * 0.6180339887498949) | 1;
        while (gcd64(mul, p.n_tiles_total) != 1) mul += 2;
    }
    p.perm_mul = mul;

    // Segment sizes: 16 tiles, then at most kSegGrowth x (tiles already seen).  With tau = the
    // KC-th best of the rows seen so far, a segment of g x seen rows appends ~ g*KC keys per query
    // (cap = 32*KC leaves an 8x margin over the expectation at g = 4).  Smaller g = fewer appends
    // (total ~ KC * g * log_{1+g}(N/4096) per query) but more launches; g = 4 gives 7 segments at 10M rows.
    static const int64_t kSegGrowth = [] {           // CADENCE_K2_GROWTH: A/B aid (2..12), default 4
        const char *e = getenv("CADENCE_K2_GROWTH");
        const int v = e ? atoi(e) : 4;
        return (int64_t)((v >= 1 && v <= 12) ? v : 4);
    }();
    // first segment (every row is appended: tau = -inf): 256 * kSeg0 rows must fit the list capacity 32 * KC >= 4096
    static const int64_t kSeg0 = [] {                // CADENCE_K2_SEG0: A/B aid (4, 8, 16), default 16
        const char *e = getenv("CADENCE_K2_SEG0");
        const int v = e ? atoi(e) : 16;
        return (int64_t)((v == 4 || v == 8 || v == 16) ? v : 16);
    }();
    // A persistent grid must be co-resident: with 4-CTA clusters a GPC's SM count need not be a multiple of 4, so the
    // number of clusters that fit can be smaller than sm_count / 4 (a cluster that waits for a free slot would run
    // its share of the segment alone afterwards).
    int64_t resident_clusters = s->sm_count / cluster;
    if (cluster == 4) {
        static std::mutex occ_mu;
        static int occ_by_dev[64] = {0};
        std::lock_guard<std::mutex> occ_lock(occ_mu);
        if (occ_by_dev[s->device & 63] == 0) {
            cudaLaunchConfig_t oc = {};
            oc.gridDim = dim3((unsigned)(s->sm_count / 4 * 4));
            oc.blockDim = dim3(gemm_threads(2));
            oc.dynamicSmemBytes = gemm_smem(2);
            cudaLaunchAttribute oa[1];
            oa[0].id = cudaLaunchAttributeClusterDimension;
            oa[0].val.clusterDim.x = 4; oa[0].val.clusterDim.y = 1; oa[0].val.clusterDim.z = 1;
            oc.attrs = oa; oc.numAttrs = 1;
            int n_act = 0;
            CDR_CUDA(cudaOccupancyMaxActiveClusters(&n_act, (const void *)fns[3][1], &oc));
            occ_by_dev[s->device & 63] = n_act > 0 ? n_act : 1;
            if (getenv("CADENCE_K2_VERBOSE")) fprintf(stderr, "K2: %d co-resident 4-CTA clusters on device %d\n", n_act, s->device);
        }
        if (occ_by_dev[s->device & 63] < resident_clusters) resident_clusters = occ_by_dev[s->device & 63];
    }
    int64_t begin = 0;
    while (begin < p.n_tiles_total) {
        int64_t end = begin == 0 ? kSeg0 : begin + kSegGrowth * begin;
        if (end > p.n_tiles_total) end = p.n_tiles_total;
        p.tile_begin = begin;
        p.tile_end = end;
        int64_t max_clusters = resident_clusters;
        // tile-major walk only where every cluster gets >= 32 corpus tiles (imbalance <= 1/32)
        p.tile_major = (k2_order == 1 && (end - begin) >= 32 * max_clusters && p.m_tiles / cluster > 1) ? 1 : 0;
        const int64_t items = p.tile_major ? (end - begin) : (end - begin) * (p.m_tiles / cluster);
        const int grid = (int)(items < max_clusters ? items : max_clusters) * cluster;
        cdr_prof_mark_begin(1, st);
        {
            const int epi = epi_forced ? epi_forced : (p.tile_major ? 1 : 2);
            const gemm_fn fn = fns[form][epi == 4 ? 2 : epi - 1];
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(grid);
            cfg.blockDim = dim3(gemm_threads(epi));
            cfg.dynamicSmemBytes = gemm_smem(epi);
            cfg.stream = st;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = cluster;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = cluster > 1 ? 1 : 0;
            CDR_CUDA(cudaLaunchKernelEx(&cfg, fn, map_q, map_x, p));
        }
        CDR_LAUNCH_CHECK();
        cdr_prof_mark_end(1, st);
        if (end < p.n_tiles_total) {
            if (kc == 128) select_compact_kernel<4><<<nq, 256, 0, st>>>(lists, counts, tau, tau_key, overflow, cap);
            else select_compact_kernel<8><<<nq, 256, 0, st>>>(lists, counts, tau, tau_key, overflow, cap);
            CDR_LAUNCH_CHECK();
        }
        begin = end;
    }

    rc = cdr_finalize_unsorted_launch(s, lists, counts, cap, kc, q_dev, nq, k, s->emb_f32 == nullptr, out_score_dev,
                                      out_id_dev, out_n_dev, st);
    if (rc != CDR_OK) return rc;

    // ---- candidate-list overflow (adversarial data): detected and repaired ON THE DEVICE, no host round trip.  A
    // one-block kernel lists the overflowed queries; the exact lane then re-runs exactly those (count read from device
    // memory by the kernels themselves: with none listed -- always, on real data -- its two launches exit at once).
    // Short results need no second look: NaN-scored rows are admitted like in the exact lane, so fewer than k results
    // means fewer than k allowed rows.
    const int rounds = (nq + kRedoSlots - 1) / kRedoSlots;
    redo_list_kernel<<<1, 1024, 0, st>>>(overflow, counts, cap, nq, redo_idx, redo_cnt, kRedoSlots, rounds);
    CDR_LAUNCH_CHECK();
    for (int r = 0; r < rounds; ++r) {
        const int n_slots = nq - r * kRedoSlots < kRedoSlots ? nq - r * kRedoSlots : kRedoSlots;
        if (s->emb_f32 != nullptr)
            rc = cdr_exact_scan_redo_launch(s, ws, q_dev, n_slots, p.allow, k, out_score_dev, out_id_dev, out_n_dev, st,
                                            redo_idx + r * kRedoSlots, redo_cnt + r);
        else      // bf16-only store: the scan lane over the bf16 rows is exact over the rows as stored
            rc = cdr_bf16_scan_launch(s, ws, q_dev, n_slots, p.allow, k, out_score_dev, out_id_dev, out_n_dev, st,
                                      redo_idx + r * kRedoSlots, redo_cnt + r);
        if (rc != CDR_OK) return rc;
    }
    return CDR_OK;
}
