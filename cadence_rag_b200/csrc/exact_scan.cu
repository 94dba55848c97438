// exact_scan.cu -- K1 (HBM-bound fp32 cosine scan with fused running top-k) and the
// K3/K4 finalize kernel (k-way merge of the per-CTA lists + fp64 re-score + final order).
//
// Replaces the reference's mode="exact" dense SQL (app/retrieve.py:339-351, 374-386):
//   ORDER BY embedding <=> :q LIMIT :k  over  WHERE <filters> AND embedding IS NOT NULL.
//
// K1 data flow (one persistent CTA per SM, 9 warps):
//   warp 8 lane 0 : producer.  For every row tile owned by this CTA it arms the stage's "full"
//                   mbarrier with the byte count and issues two 1-D bulk async copies
//                   (cp.async.bulk, SASS UBLKCP): TR rows (TR*dim*4 bytes, 64 KB at dim 1024)
//                   and the TR inverse norms.  Stages form a ring; a stage is re-armed when the
//                   8 consumer warps have arrived on its "empty" mbarrier.
//   warps 0..7    : consumers.  The query lives in registers (J float4 per lane).  Each warp
//                   owns RPW rows of the tile, reads them from shared memory as conflict-free
//                   128-bit loads, accumulates the dot product in fp32, butterfly-reduces it,
//                   scales by 1/||x|| * 1/||q|| and pushes (score,row) keys that beat the warp's
//                   current KC-th best into a per-warp list in shared memory.
//   epilogue      : each warp sorts its list (register bitonic network), the 8 lists are
//                   merged pairwise through shared memory, and the CTA writes its KC best keys.
// The corpus is read exactly once per query; nothing but KC keys per CTA goes back to HBM.
// Variants of the same kernel (template parameters QPC / BF, see ScanSmem): shared reads for batches (3 queries in
// registers; 16 queries per CTA as register tiles), the gather launch for selective filters, and the candidate pass
// of the single-query "ann" lane over the bf16 rows.
//
// Algorithmic bytes per query: n_rows * dim * 4.
#include "common.cuh"

#include <cooperative_groups.h>
#include <cstdlib>
#include <mutex>

#ifndef CDR_BF16_FHFMA
#define CDR_BF16_FHFMA 1      // 0: the widening (shift / mask + FFMA) form of the bf16 scan, for A/B builds
#endif

namespace {

// two bf16 values (RN-even) in one register, first value in the low half
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi)
{
    const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t *>(&v);
}
// acc += row.lo * query.lo + row.hi * query.hi, fp32 accumulate, low half first (two FHFMA.BF16)
__device__ __forceinline__ float fhfma_bf16x2(uint32_t row, uint32_t query, float acc)
{
    asm("{\n"
        ".reg .b16 rl, rh, ql, qh;\n"
        "mov.b32 {rl, rh}, %1;\n"
        "mov.b32 {ql, qh}, %2;\n"
        "fma.rn.f32.bf16 %0, rl, ql, %0;\n"
        "fma.rn.f32.bf16 %0, rh, qh, %0;\n"
        "}\n"
        : "+f"(acc)
        : "r"(row), "r"(query));
    return acc;
}

constexpr int kConsumerWarps = 8;
constexpr int kMaxStages = 6;

struct ScanParams {
    const float *rows;        // [n_rows, dim]
    const float *inv_norm;    // [>= round_up(n_rows, TR)]
    const uint32_t *allow;    // nullable bitmap
    const float *queries;     // [nq, dim]
    int nq;                   // queries of this launch; gridDim.y = ceil(nq / QPC) query groups
    uint64_t *cta_keys;       // [nq, lists_per_query, KC]
    int64_t n_rows;
    int64_t n_tiles;
    int n_stages;
    unsigned int *tile_ctr;   // [nq] zero on entry (reset by the finalize kernel); nullptr = static tile assignment
    // selective filters (see launch_scan_t): the allowed rows as a compact list + their count on the device.
    //   row_list != nullptr : GATHER launch -- scans list entries [0, *list_count); does nothing when the
    //                         count exceeds list_cap (the list is then truncated and the full scan serves)
    //   row_list == nullptr && list_count != nullptr : FULL launch that does nothing when the gather launch serves
    const uint32_t *row_list;
    const unsigned int *list_count;
    uint32_t list_cap;
    int lists_per_query;      // per-query stride of cta_keys in lists (gather CTAs + full CTAs)
    int list_offset;          // first list of this launch inside a query's block
    // conditional re-run of a few queries of a batch (K2's overflow path, QPC == 1 only): the launch serves
    // min(*q_count, nq) SLOTS, slot g scans for query q_index[g] and writes the lists of slot g; with a zero count
    // every CTA exits at once.  Both null = the plain launch (slot == query, nq of them).
    const int *q_index;
    const int *q_count;
};

// Dynamic shared memory carve-up (computed identically on host and device).
// QPC = queries per CTA: 1 = one scan of the corpus per query (the single-query GEMV of BASELINE configs[1]);
// kSharedQPC = "shared reads": a CTA scores every tile it streams against 3 queries held in registers, so a
// batch of concurrent exact requests costs a third of the HBM traffic.  Per-query arithmetic is identical (same
// lanes, same accumulation order), so both give the same bits.  Why 3: with 9 warps one SM sub-partition hosts
// 3 warps, i.e. <= 170 registers per thread; 4 queries (128 registers of query data) spill, and because the CTA
// takes ~all of the unified L1/shared memory for its tile ring the spills go to L2 (measured: 1.26x instead of
// 4x).  3 queries compile to 168 registers with no spills.
//
// kDeepQPC = "deep shared reads" for whole groups of 16 exact requests (and tails of >= 10): a register-tiled
// kernel.  The 16 queries of the CTA are staged once in shared memory (64 KB); the 8 consumer warps form 4 row groups
// x 2 query halves, and a warp keeps the 4 rows x 8 queries x 2 partial sums of its part of the tile in registers
// while it streams both operands from shared memory: 4 + 8 LDS.128 feed 4 x 8 x 4 FMAs per 128-byte column step (12
// shared-memory wavefronts per (row, query)).  The top-k state of a warp's 8 queries lives in shared memory too
// (thresholds are read back per candidate, pushes are rare), the tile ring shrinks to 2 stages, 168 registers, no
// spills.  Same lanes, same FMA order, same reduction => same bits as one scan per query.
constexpr int kSharedQPC = 3;
constexpr int kDeepQPC = 16;
// Shape of the deep kernel: 8 consumer warps as 4 row groups x kDeepQSplit query halves; a warp owns 4 rows of the
// 16-row tile and 8 of the CTA's 16 queries.  (Rows resident in registers with the queries streamed past them -- 8 warps
// x 2 rows, or 4 warps x 4 rows -- cost 20 / 12 LDS wavefronts per (row, query) but left the LSU pipe 71 % busy and the
// 4-warp form latency-bound: 7.4 / 8.4 ms per 64 queries over 1 M rows, profiles/r01/k1_shared_probe_deep4w.json.)
constexpr int kDeepWarps = 8;
constexpr int kDeepRPW = 4;
constexpr int kDeepQSplit = 2;
struct DeepTopK {     // per (consumer warp, query) in shared memory
    uint64_t tau;
    int count;
    int min_pos;
};
// BF = the rows streamed are the store's bf16 copy (row L2-normalised, RN-even) instead of the fp32 rows: the
// candidate pass of the single-query "ann" lane (cdr_search_scan_bf16), half the bytes per row.
template <int J, int RPW, int NPL, int QPC, bool BF = false>
struct ScanSmem {
    static constexpr int DIM = J * 128;
    static constexpr size_t kRowBytes = (size_t)DIM * (BF ? 2 : 4);
    static constexpr bool kDeep = !BF && QPC > kSharedQPC;              // (bf16-row passes shared by 2 or 4 queries keep them in registers)
    static constexpr int CW = kDeep ? kDeepWarps : kConsumerWarps;              // consumer warps
    static constexpr int QW = kDeep ? QPC / kDeepQSplit : QPC;                   // queries a warp keeps top-k lists for
    static constexpr int TR = kDeep ? RPW * CW / kDeepQSplit : RPW * CW;
    static constexpr int KC = NPL * 32;
    static constexpr size_t kTileBytes = (size_t)TR * kRowBytes;
    static constexpr size_t kMetaBytes = (size_t)TR * 4;   // inverse norms (multiple of 16)
    static constexpr size_t kListBytes = (size_t)CW * QW * KC * 8;
    static constexpr size_t kBarBytes = 3 * kMaxStages * 8;   // full + empty barriers + the tile index of each stage
    // deep shared reads (QPC > kSharedQPC): the queries, their inverse norms and the per-(warp, query) top-k
    // state live in shared memory instead of registers
    static constexpr size_t kQueryBytes = kDeep ? (size_t)QPC * DIM * 4 : 0;
    static constexpr size_t kStateBytes = kDeep ? (size_t)CW * QW * 16 + (size_t)QPC * 4 : 0;
    static constexpr size_t bytes(int stages)
    {
        return (size_t)stages * (kTileBytes + kMetaBytes) + kQueryBytes + kListBytes + kStateBytes + kBarBytes + 128;
    }
    static int max_stages(size_t smem_limit)
    {
        int s = kMaxStages;
        while (s > 1 && bytes(s) > smem_limit) --s;
        return s;
    }
};

template <int J, int RPW, int NPL, int QPC, bool BF = false>
__global__ void __launch_bounds__((ScanSmem<J, RPW, NPL, QPC, BF>::CW + 1) * 32, 1) exact_scan_kernel(const ScanParams p)
{
    using L = ScanSmem<J, RPW, NPL, QPC, BF>;
    constexpr int DIM = L::DIM, TR = L::TR, KC = L::KC, CW = L::CW;
    extern __shared__ __align__(128) unsigned char smem_raw[];

    // Programmatic dependent launch: the finalize kernel behind this launch may be scheduled now (its CTAs fit beside a
    // scan CTA in the latency form) and sleeps in griddepcontrol.wait until this grid has completed and its lists are
    // visible -- its launch latency and query-norm prologue leave the critical path.  No effect on a plain launch.
    asm volatile("griddepcontrol.launch_dependents;");

    const int S = p.n_stages;
    unsigned char *tiles = smem_raw;
    unsigned char *metas = tiles + (size_t)S * L::kTileBytes;
    float4 *qs = reinterpret_cast<float4 *>(metas + (size_t)S * L::kMetaBytes);           // deep: [QPC][DIM/4]
    uint64_t *lists = reinterpret_cast<uint64_t *>(reinterpret_cast<unsigned char *>(qs) + L::kQueryBytes);
    DeepTopK *dstate = reinterpret_cast<DeepTopK *>(lists + CW * L::QW * KC);   // deep: [warps][QW]
    float *qinv_s = reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(dstate) + (size_t)CW * L::QW * 16 * (L::kDeep ? 1 : 0));
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(reinterpret_cast<unsigned char *>(dstate) + L::kStateBytes);
    uint64_t *empty_bar = full_bar + kMaxStages;
    long long *stage_tile = reinterpret_cast<long long *>(empty_bar + kMaxStages);   // tile held by a stage, -1 = end

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // A CTA is persistent over query groups: group g = blockIdx.y, blockIdx.y + gridDim.y, ...  The stage ring
    // (counter `it` below, advanced identically by the producer and the consumers, end-of-stream stages included)
    // runs on across groups, so the producer prefetches the next group's first tiles while the consumers sort and
    // merge the lists of the current one.
    int nq_eff = p.nq;
    if (p.q_count != nullptr) {
        const int c = *p.q_count;                     // written earlier on the stream
        nq_eff = c < p.nq ? c : p.nq;
    }
    const int n_groups = (nq_eff + QPC - 1) / QPC;
    if ((int)blockIdx.y >= n_groups) return;
    const int64_t G = gridDim.x;
    const bool gather = p.row_list != nullptr;
    int64_t n_tiles = p.n_tiles;
    uint32_t n_listed = 0;
    if (p.list_count != nullptr) {
        n_listed = *p.list_count;                     // written by compact_allow_kernel earlier on the stream
        const bool list_serves = n_listed <= p.list_cap;
        if (gather != list_serves) {                  // the other launch of the pair does the work: empty lists
            if (warp == 0) {
                for (int g = blockIdx.y; g < n_groups; g += gridDim.y) {
                    const int gq0 = g * QPC;
                    const int gn = nq_eff - gq0 < QPC ? nq_eff - gq0 : QPC;
                    for (int u = 0; u < gn; ++u) {
                        uint64_t *out = p.cta_keys + ((size_t)(gq0 + u) * p.lists_per_query + p.list_offset + blockIdx.x) * KC;
#pragma unroll
                        for (int i = 0; i < NPL; ++i) out[i * 32 + lane] = CDR_EMPTY_KEY;
                    }
                }
            }
            return;
        }
        if (gather) n_tiles = ((int64_t)n_listed + TR - 1) / TR;
    }

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], CW);
        }
        fence_mbar_init();
    }
    __syncthreads();

    // Tile assignment.  The first S tiles of a CTA are static (blockIdx.x + i*G: no latency at start-up);
    // after that the producer draws tiles from a per-query counter (work stealing), so SMs that sit
    // closer to the memory partitions simply scan more tiles and all CTAs finish together -- with a
    // static split the slowest SM sets the latency of a single-query scan.  tile_ctr == nullptr keeps
    // the static split (t = blockIdx.x + i*G).  A stage whose tile index is -1 ends the stream.
    const bool dynamic = p.tile_ctr != nullptr;

    if (warp == CW && gather) {
        // ------------------------------------------------------------------ producer, gather launch
        // The whole warp runs the loop: lane 0 draws tiles and arms the barrier, lanes 0..TR-1 each fetch one
        // list entry and issue that row's bulk copy, so the TR copies of a tile are issued in parallel.
        static_assert(TR <= 32, "one lane per row of a tile");
        uint32_t it = 0;
        for (int g = blockIdx.y; g < n_groups; g += gridDim.y) {
        int64_t next_dyn = -1;
        for (int64_t i = 0;; ++i, ++it) {
            const int s = (int)(it % (uint32_t)S);
            const uint32_t ph = (it / (uint32_t)S) & 1u;
            int64_t tile = blockIdx.x + i * G;
            if (dynamic && i >= S) tile = next_dyn;
            if (dynamic && i + 1 >= S) {
                unsigned int t = 0;
                if (lane == 0) t = atomicAdd(&p.tile_ctr[g], 1u);
                next_dyn = (int64_t)S * G + __shfl_sync(0xffffffffu, t, 0);
            }
            // the list entry of this lane's row can be fetched while the stage is still busy
            const int64_t e0 = tile * TR;
            const int nr = tile < n_tiles ? (int)((int64_t)n_listed - e0 < TR ? (int64_t)n_listed - e0 : TR) : 0;
            const uint32_t my_row = lane < nr ? __ldg(&p.row_list[e0 + lane]) : 0u;
            if (lane == 0) mbar_wait(&empty_bar[s], ph ^ 1u);
            __syncwarp();
            if (tile >= n_tiles) {
                if (lane == 0) {
                    stage_tile[s] = -1;
                    mbar_arrive(&full_bar[s]);
                }
                ++it;
                break;
            }
            if (lane == 0) {
                stage_tile[s] = tile;
                mbar_arrive_expect_tx(&full_bar[s], (uint32_t)(nr * L::kRowBytes));
            }
            __syncwarp();
            if (lane < nr)
                bulk_g2s(tiles + (size_t)s * L::kTileBytes + (size_t)lane * L::kRowBytes,
                         reinterpret_cast<const unsigned char *>(p.rows) + (size_t)my_row * L::kRowBytes,
                         (uint32_t)L::kRowBytes, &full_bar[s]);
        }
        }
    } else if (warp == CW) {
        // ------------------------------------------------------------------ producer
        if (lane == 0) {
            uint32_t it = 0;
            for (int g = blockIdx.y; g < n_groups; g += gridDim.y) {
            int64_t next_dyn = -1;
            for (int64_t i = 0;; ++i, ++it) {
                const int s = (int)(it % (uint32_t)S);
                const uint32_t ph = (it / (uint32_t)S) & 1u;
                int64_t tile = blockIdx.x + i * G;
                if (dynamic && i >= S) tile = next_dyn;
                // draw the next tile now: the atomic's round trip overlaps the wait for a free stage
                if (dynamic && i + 1 >= S) next_dyn = (int64_t)S * G + atomicAdd(&p.tile_ctr[g], 1u);
                mbar_wait(&empty_bar[s], ph ^ 1u);
                if (tile >= n_tiles) {
                    stage_tile[s] = -1;
                    mbar_arrive(&full_bar[s]);
                    ++it;
                    break;
                }
                stage_tile[s] = tile;
                const int64_t row0 = tile * TR;
                int64_t nr = p.n_rows - row0;
                if (nr > TR) nr = TR;
                const uint32_t row_bytes = (uint32_t)(nr * L::kRowBytes);
                mbar_arrive_expect_tx(&full_bar[s], row_bytes + (uint32_t)L::kMetaBytes);
                bulk_g2s(tiles + (size_t)s * L::kTileBytes,
                         reinterpret_cast<const unsigned char *>(p.rows) + (size_t)row0 * L::kRowBytes, row_bytes,
                         &full_bar[s]);
                bulk_g2s(metas + (size_t)s * L::kMetaBytes, p.inv_norm + row0,
                         (uint32_t)L::kMetaBytes, &full_bar[s]);
            }
            }
        }
    } else {
        // ------------------------------------------------------------------ consumers
        uint32_t it = 0;
        for (int g = blockIdx.y; g < n_groups; g += gridDim.y) {
        const int q0 = g * QPC;                                           // first SLOT of this group (lists are per slot)
        const int nqv = nq_eff - q0 < QPC ? nq_eff - q0 : QPC;            // valid queries in the group
        const int qsrc0 = (QPC == 1 && p.q_index != nullptr) ? __ldg(&p.q_index[q0]) : q0;   // first QUERY of the group
        if constexpr (L::kDeep) {
            // ---- deep shared reads: a 4-row x 8-query register tile per warp, rows and queries streamed from shared memory
            constexpr int QW = L::QW;                         // queries of this warp (8)
            constexpr int GQ = 8 / RPW;                       // queries per joint reduction (8 values)
            static_assert(RPW * GQ == 8 && QW % GQ == 0, "deep shared reads: RPW must divide 8");
            const int rg = warp / kDeepQSplit;                // row group: rows rg*RPW .. of the tile
            const int qh = warp % kDeepQSplit;                // query half: queries qh*QW .. of the CTA's group
            for (int u = warp; u < QPC; u += CW) {
                float qn = 0.f;
                const float4 *qv = reinterpret_cast<const float4 *>(p.queries + (size_t)(qsrc0 + (u < nqv ? u : 0)) * DIM);
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    const float4 t = __ldg(&qv[j * 32 + lane]);
                    qs[u * (DIM / 4) + j * 32 + lane] = t;
                    qn = fmaf(t.x, t.x, qn);
                    qn = fmaf(t.y, t.y, qn);
                    qn = fmaf(t.z, t.z, qn);
                    qn = fmaf(t.w, t.w, qn);
                }
                qn = warp_sum_f32(qn);
                if (lane == 0) qinv_s[u] = __fdiv_rn(1.0f, __fsqrt_rn(qn));
            }
            DeepTopK *my_state = dstate + warp * QW;
            uint64_t *my_lists = lists + (size_t)warp * QW * KC;
            for (int i = lane; i < QW * KC; i += 32) my_lists[i] = CDR_EMPTY_KEY;
            if (lane < QW) {
                my_state[lane].tau = 0;
                my_state[lane].count = 0;
                my_state[lane].min_pos = 0;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(CW * 32) : "memory");
            const float4 *my_q = qs + (size_t)(qh * QW) * (DIM / 4) + lane;
            const float *my_qinv = qinv_s + qh * QW;

            for (;; ++it) {
                const int s = (int)(it % (uint32_t)S);
                const uint32_t ph = (it / (uint32_t)S) & 1u;
                mbar_wait(&full_bar[s], ph);
                const int64_t tile = stage_tile[s];
                if (tile < 0) {                                   // end of this group's stream: hand the stage back
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty_bar[s]);
                    ++it;
                    break;
                }
                CDR_DEV_ASSERT(s >= 0 && s < S && S <= kMaxStages && tile < n_tiles);
                const int64_t row0 = tile * TR + rg * RPW;    // a multiple of RPW: the rows share one bitmap word
                uint32_t allow_bits = 0xFFFFFFFFu;
                uint32_t g_row[RPW];
                float g_inv[RPW];
                if (gather) {
#pragma unroll
                    for (int r = 0; r < RPW; ++r) {
                        const bool in = row0 + r < (int64_t)n_listed;
                        g_row[r] = in ? __ldg(&p.row_list[row0 + r]) : 0u;
                        g_inv[r] = __ldg(&p.inv_norm[g_row[r]]);
                        if (!in) allow_bits &= ~(1u << r);
                    }
                } else if (p.allow != nullptr) {
                    const uint32_t w = (row0 < p.n_rows) ? __ldg(&p.allow[row0 >> 5]) : 0u;
                    allow_bits = w >> (row0 & 31);
                }
                const float4 *tv = reinterpret_cast<const float4 *>(tiles + (size_t)s * L::kTileBytes) +
                                   (size_t)(rg * RPW) * (DIM / 4) + lane;
                const float *mv = reinterpret_cast<const float *>(metas + (size_t)s * L::kMetaBytes) + rg * RPW;
                float acc[QW][RPW][2];
#pragma unroll
                for (int u = 0; u < QW; ++u)
#pragma unroll
                    for (int r = 0; r < RPW; ++r) acc[u][r][0] = acc[u][r][1] = 0.f;
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    float4 rv[RPW];
#pragma unroll
                    for (int r = 0; r < RPW; ++r) rv[r] = tv[r * (DIM / 4) + j * 32];
#pragma unroll
                    for (int u = 0; u < QW; ++u) {
                        const float4 qv = my_q[u * (DIM / 4) + j * 32];   // RPW + QW LDS.128 feed RPW x QW x 4 FMAs
#pragma unroll
                        for (int r = 0; r < RPW; ++r) {
                            float a = acc[u][r][j & 1];
                            a = fmaf(rv[r].x, qv.x, a);
                            a = fmaf(rv[r].y, qv.y, a);
                            a = fmaf(rv[r].z, qv.z, a);
                            a = fmaf(rv[r].w, qv.w, a);
                            acc[u][r][j & 1] = a;
                        }
                    }
                }
                float inv_n[RPW];
#pragma unroll
                for (int r = 0; r < RPW; ++r) inv_n[r] = gather ? g_inv[r] : mv[r];
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_bar[s]);        // all reads of the stage are done

                const int v = (lane >> 2) & 7;                    // the value this 4-lane group owns after a reduction
                const int vu = v / RPW, vr = v % RPW;
                float s_inv_n = inv_n[0];
                int64_t s_row = gather ? (int64_t)g_row[0] : row0;
#pragma unroll
                for (int r = 1; r < RPW; ++r)
                    if (vr == r) { s_inv_n = inv_n[r]; s_row = gather ? (int64_t)g_row[r] : row0 + r; }
                const bool s_ok = (s_row < p.n_rows) && ((allow_bits >> vr) & 1u);
                const bool b4 = (lane & 16) != 0, b3 = (lane & 8) != 0, b2 = (lane & 4) != 0;
#pragma unroll
                for (int g = 0; g < QW / GQ; ++g) {
                    float a8[8];
#pragma unroll
                    for (int x = 0; x < 8; ++x) a8[x] = acc[g * GQ + x / RPW][x % RPW][0] + acc[g * GQ + x / RPW][x % RPW][1];
                    // the transposed reduction of the QPC <= 3 path (same additions per lane as the butterfly)
                    float h4[4], h2[2];
#pragma unroll
                    for (int i2 = 0; i2 < 4; ++i2)
                        h4[i2] = (b4 ? a8[i2 + 4] : a8[i2]) + __shfl_xor_sync(0xffffffffu, b4 ? a8[i2] : a8[i2 + 4], 16);
#pragma unroll
                    for (int i2 = 0; i2 < 2; ++i2)
                        h2[i2] = (b3 ? h4[i2 + 2] : h4[i2]) + __shfl_xor_sync(0xffffffffu, b3 ? h4[i2] : h4[i2 + 2], 8);
                    float dot = (b2 ? h2[1] : h2[0]) + __shfl_xor_sync(0xffffffffu, b2 ? h2[0] : h2[1], 4);
                    dot += __shfl_xor_sync(0xffffffffu, dot, 2);
                    dot += __shfl_xor_sync(0xffffffffu, dot, 1);
                    const int u_mine = g * GQ + vu;
                    const uint64_t key = cdr_pack_key(dot * s_inv_n * my_qinv[u_mine], (uint32_t)s_row);
                    unsigned pend = __ballot_sync(0xffffffffu, s_ok && key > my_state[u_mine].tau) & 0x11111111u;
                    while (pend) {                                // rare; warp-uniform
                        const int src = __ffs(pend) - 1;
                        pend &= pend - 1;
                        const uint64_t k1 = __shfl_sync(0xffffffffu, key, src);
                        const int u1 = g * GQ + (src >> 2) / RPW;
                        DeepTopK *ts = my_state + u1;
                        uint64_t *tl = my_lists + (size_t)u1 * KC;
                        if (k1 > ts->tau) {                       // re-tested: an earlier insert may have raised tau
                            const int cnt = __shfl_sync(0xffffffffu, ts->count, 0);   // lane 0 is the only writer
                            CDR_DEV_ASSERT(cnt >= 0 && cnt <= KC && ts->min_pos >= 0 && ts->min_pos < KC);
                            if (lane == 0) {
                                tl[cnt < KC ? cnt : ts->min_pos] = k1;
                                if (cnt < KC) ts->count = cnt + 1;
                            }
                            __syncwarp();
                            if (cnt + 1 >= KC) {                  // list full: new threshold = its minimum
                                uint64_t m = ~0ull;
                                int pos = 0;
#pragma unroll
                                for (int i2 = 0; i2 < NPL; ++i2) {
                                    const uint64_t x = tl[i2 * 32 + lane];
                                    if (x < m) { m = x; pos = i2 * 32 + lane; }
                                }
#pragma unroll
                                for (int o = 16; o > 0; o >>= 1) {
                                    const uint64_t om = __shfl_xor_sync(0xffffffffu, m, o);
                                    const int op = __shfl_xor_sync(0xffffffffu, pos, o);
                                    if (om < m) { m = om; pos = op; }
                                }
                                if (lane == 0) { ts->tau = m; ts->min_pos = pos; }
                            }
                            __syncwarp();
                        }
                    }
                }
            }

            // ---- per query of this warp: sort, then merge across the row groups of the same query half
#pragma unroll 1
            for (int u = 0; u < QW; ++u) {
                uint64_t *mine = my_lists + (size_t)u * KC;
                uint64_t k[NPL];
                __syncwarp();
#pragma unroll
                for (int i = 0; i < NPL; ++i) k[i] = mine[i * 32 + lane];
                warp_bitonic_sort_desc<NPL>(k, lane);
#pragma unroll
                for (int i = 0; i < NPL; ++i) mine[i * 32 + lane] = k[i];
#pragma unroll
                for (int step = 1; step < CW / kDeepQSplit; step <<= 1) {
                    asm volatile("bar.sync 1, %0;" ::"n"(CW * 32) : "memory");
                    if ((rg & (2 * step - 1)) == 0)
                        warp_merge_topk<NPL>(k, lists + ((size_t)(warp + step * kDeepQSplit) * QW + u) * KC, lane);
                    asm volatile("bar.sync 1, %0;" ::"n"(CW * 32) : "memory");
                    if ((rg & (2 * step - 1)) == 0) {
#pragma unroll
                        for (int i = 0; i < NPL; ++i) mine[i * 32 + lane] = k[i];
                    }
                }
                const int uq = qh * QW + u;
                if (rg == 0 && uq < nqv) {
                    CDR_DEV_ASSERT(q0 + uq < nq_eff && p.list_offset + (int)blockIdx.x < p.lists_per_query);
                    uint64_t *out = p.cta_keys + ((size_t)(q0 + uq) * p.lists_per_query + p.list_offset + blockIdx.x) * KC;
#pragma unroll
                    for (int i = 0; i < NPL; ++i) out[i * 32 + lane] = k[i];
                }
            }
        } else if constexpr (BF) {
            // ---- candidate pass over the bf16 rows (the "ann" scan lane): a lane takes 8 consecutive elements
            // per 16-byte step and accumulates in fp32 against the query.  The rows are stored normalised, so the
            // score is dot / |q|; survivors are re-scored exactly (fp64 on the fp32 rows) by the finalize kernel.
            // QPC == 1: one query per pass, 4 rows per warp.  QPC x RPW == 8 (built: 2 x 4): the queries of a batch
            // share passes over the rows -- every 16-byte shared-memory load feeds QPC queries.
            static_assert(J % 2 == 0, "bf16 scan: dim a multiple of 256");
            static_assert(QPC == 1 || QPC * RPW == 8, "bf16 scan: shared passes reduce 8 dot products together");
            static_assert(QPC == 1 || CDR_BF16_FHFMA, "bf16 scan: shared passes use the mixed-precision FMA form");
            constexpr int JB = J / 2;
            float inv_q[QPC];
            uint32_t qh[QPC][JB][4];
            float4 qa[QPC == 1 ? JB : 1], qb[QPC == 1 ? JB : 1];       // fp32 query of the widening form (A/B build)
            (void)qa; (void)qb;
            WarpTopK<NPL> top[QPC];
#pragma unroll
            for (int u = 0; u < QPC; ++u) {
                float qn = 0.f;
                // a missing query of the last group re-reads the group's first query; its results are never written
                const float4 *qv = reinterpret_cast<const float4 *>(p.queries + (size_t)(qsrc0 + (u < nqv ? u : 0)) * DIM);
#pragma unroll
                for (int j = 0; j < JB; ++j) {
                    const float4 xa = __ldg(&qv[(j * 32 + lane) * 2]);
                    const float4 xb = __ldg(&qv[(j * 32 + lane) * 2 + 1]);
                    qn = fmaf(xa.x, xa.x, qn); qn = fmaf(xa.y, xa.y, qn);
                    qn = fmaf(xa.z, xa.z, qn); qn = fmaf(xa.w, xa.w, qn);
                    qn = fmaf(xb.x, xb.x, qn); qn = fmaf(xb.y, xb.y, qn);
                    qn = fmaf(xb.z, xb.z, qn); qn = fmaf(xb.w, xb.w, qn);
                    // The query as bf16 pairs (RN-even, like the tensor-core lane's): a row element and its query element
                    // then meet in ONE mixed-precision FMA (fma.rn.f32.bf16, SASS FHFMA.BF16 with .H0/.H1 operand
                    // selectors, fp32 accumulate) -- 8 instructions per 16-byte shared-memory load instead of 8 shift/mask +
                    // 8 FFMA, which left this scan issue-bound at 5.9-6.7 TB/s.  Candidate scores only.
                    qh[u][j][0] = pack_bf16x2(xa.x, xa.y); qh[u][j][1] = pack_bf16x2(xa.z, xa.w);
                    qh[u][j][2] = pack_bf16x2(xb.x, xb.y); qh[u][j][3] = pack_bf16x2(xb.z, xb.w);
                    if constexpr (QPC == 1) { qa[j] = xa; qb[j] = xb; }
                }
                qn = warp_sum_f32(qn);
                inv_q[u] = __fdiv_rn(1.0f, __fsqrt_rn(qn));
                top[u].init(lists + (warp * QPC + u) * KC, lane);
            }
            for (;; ++it) {
                const int s = (int)(it % (uint32_t)S);
                const uint32_t ph = (it / (uint32_t)S) & 1u;
                mbar_wait(&full_bar[s], ph);
                const int64_t tile = stage_tile[s];
                if (tile < 0) {                                   // end of this group's stream: hand the stage back
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty_bar[s]);
                    ++it;
                    break;
                }
                CDR_DEV_ASSERT(s >= 0 && s < S && S <= kMaxStages && tile < n_tiles);
                const int64_t row0 = tile * TR + warp * RPW;      // a multiple of RPW (<= 4): one bitmap word
                uint32_t allow_bits = 0xFFFFFFFFu;
                uint32_t g_row[RPW];
                if (gather) {
#pragma unroll
                    for (int r = 0; r < RPW; ++r) {
                        const bool in = row0 + r < (int64_t)n_listed;
                        g_row[r] = in ? __ldg(&p.row_list[row0 + r]) : 0u;
                        if (!in) allow_bits &= ~(1u << r);
                    }
                } else if (p.allow != nullptr) {
                    const uint32_t w = (row0 < p.n_rows) ? __ldg(&p.allow[row0 >> 5]) : 0u;
                    allow_bits = w >> (row0 & 31);
                }
                const uint4 *tv = reinterpret_cast<const uint4 *>(tiles + (size_t)s * L::kTileBytes) +
                                  (size_t)(warp * RPW) * (DIM / 8) + lane;
                float acc[QPC][RPW][2];
#pragma unroll
                for (int u = 0; u < QPC; ++u)
#pragma unroll
                    for (int r = 0; r < RPW; ++r) acc[u][r][0] = acc[u][r][1] = 0.f;
#pragma unroll
                for (int j = 0; j < JB; ++j) {
#pragma unroll
                    for (int r = 0; r < RPW; ++r) {
                        const uint4 v = tv[r * (DIM / 8) + j * 32];
#if CDR_BF16_FHFMA
#pragma unroll
                        for (int u = 0; u < QPC; ++u) {
                            float a = acc[u][r][j & 1];
                            a = fhfma_bf16x2(v.x, qh[u][j][0], a);
                            a = fhfma_bf16x2(v.y, qh[u][j][1], a);
                            a = fhfma_bf16x2(v.z, qh[u][j][2], a);
                            a = fhfma_bf16x2(v.w, qh[u][j][3], a);
                            acc[u][r][j & 1] = a;
                        }
#else
                        float a = acc[0][r][j & 1];
                        a = fmaf(__uint_as_float(v.x << 16), qa[j].x, a);
                        a = fmaf(__uint_as_float(v.x & 0xFFFF0000u), qa[j].y, a);
                        a = fmaf(__uint_as_float(v.y << 16), qa[j].z, a);
                        a = fmaf(__uint_as_float(v.y & 0xFFFF0000u), qa[j].w, a);
                        a = fmaf(__uint_as_float(v.z << 16), qb[j].x, a);
                        a = fmaf(__uint_as_float(v.z & 0xFFFF0000u), qb[j].y, a);
                        a = fmaf(__uint_as_float(v.w << 16), qb[j].z, a);
                        a = fmaf(__uint_as_float(v.w & 0xFFFF0000u), qb[j].w, a);
                        acc[0][r][j & 1] = a;
#endif
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_bar[s]);
                if constexpr (QPC > 1) {
                    // QPC x RPW = 8 dot products reduced TOGETHER, exactly as in the fp32 shared-read path below: after
                    // the exchanges at offsets 16, 8, 4 the four lanes of group g = lane / 4 own value g = u * RPW + r
                    // (same additions per value as the plain butterfly => the bits of a one-query pass)
                    float a8[8];
#pragma unroll
                    for (int v = 0; v < 8; ++v) a8[v] = acc[v / RPW][v % RPW][0] + acc[v / RPW][v % RPW][1];
                    const bool b4 = (lane & 16) != 0, b3 = (lane & 8) != 0, b2 = (lane & 4) != 0;
                    float h4[4], h2[2];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        h4[i] = (b4 ? a8[i + 4] : a8[i]) + __shfl_xor_sync(0xffffffffu, b4 ? a8[i] : a8[i + 4], 16);
#pragma unroll
                    for (int i = 0; i < 2; ++i)
                        h2[i] = (b3 ? h4[i + 2] : h4[i]) + __shfl_xor_sync(0xffffffffu, b3 ? h4[i] : h4[i + 2], 8);
                    float dot = (b2 ? h2[1] : h2[0]) + __shfl_xor_sync(0xffffffffu, b2 ? h2[0] : h2[1], 4);
                    dot += __shfl_xor_sync(0xffffffffu, dot, 2);
                    dot += __shfl_xor_sync(0xffffffffu, dot, 1);
                    const int v = (lane >> 2) & 7;
                    const int vu = v / RPW, vr = v % RPW;
                    float s_inv_q = inv_q[0];
                    uint64_t s_tau = top[0].tau;
                    int64_t s_row = gather ? (int64_t)g_row[0] : row0;
#pragma unroll
                    for (int r = 1; r < RPW; ++r)
                        if (vr == r) s_row = gather ? (int64_t)g_row[r] : row0 + r;
#pragma unroll
                    for (int u = 1; u < QPC; ++u)
                        if (vu == u) { s_inv_q = inv_q[u]; s_tau = top[u].tau; }
                    const bool s_ok = (s_row < p.n_rows) && ((allow_bits >> vr) & 1u);
                    const uint64_t key = cdr_pack_key(dot * s_inv_q, (uint32_t)s_row);
                    unsigned pend = __ballot_sync(0xffffffffu, s_ok && key > s_tau) & 0x11111111u;
                    while (pend) {                             // rare; warp-uniform
                        const int src = __ffs(pend) - 1;
                        pend &= pend - 1;
                        const uint64_t k1 = __shfl_sync(0xffffffffu, key, src);
                        const int su = (src >> 2) / RPW;
#pragma unroll
                        for (int u = 0; u < QPC; ++u)
                            if (su == u && k1 > top[u].tau) top[u].push(k1, lane);   // re-tested: an earlier insert may have raised tau
                    }
                } else if constexpr (RPW == 4) {
                    // the 4 dot products of the warp are reduced TOGETHER (as in the shared-read path): at offsets 16 and
                    // 8 a lane keeps half of its values and hands the rest to its partner, so the eight lanes of group
                    // g = lane / 8 end up owning value g -- 6 shuffles instead of 20; the group then packs and tests its
                    // candidate in parallel and the rare survivors are inserted one by one
                    float a4[4];
#pragma unroll
                    for (int r = 0; r < 4; ++r) a4[r] = acc[0][r][0] + acc[0][r][1];
                    const bool b4 = (lane & 16) != 0, b3 = (lane & 8) != 0;
                    float h2[2];
#pragma unroll
                    for (int i = 0; i < 2; ++i)
                        h2[i] = (b4 ? a4[i + 2] : a4[i]) + __shfl_xor_sync(0xffffffffu, b4 ? a4[i] : a4[i + 2], 16);
                    float dot = (b3 ? h2[1] : h2[0]) + __shfl_xor_sync(0xffffffffu, b3 ? h2[0] : h2[1], 8);
                    dot += __shfl_xor_sync(0xffffffffu, dot, 4);
                    dot += __shfl_xor_sync(0xffffffffu, dot, 2);
                    dot += __shfl_xor_sync(0xffffffffu, dot, 1);
                    const int vr = lane >> 3;                  // value (row) this 8-lane group owns: bit4 -> +2, bit3 -> +1
                    int64_t s_row = gather ? (int64_t)g_row[0] : row0;
#pragma unroll
                    for (int r = 1; r < 4; ++r)
                        if (vr == r) s_row = gather ? (int64_t)g_row[r] : row0 + r;
                    const bool s_ok = (s_row < p.n_rows) && ((allow_bits >> vr) & 1u);
                    const uint64_t key = cdr_pack_key(dot * inv_q[0], (uint32_t)s_row);
                    unsigned pend = __ballot_sync(0xffffffffu, s_ok && key > top[0].tau) & 0x01010101u;
                    while (pend) {                             // rare; warp-uniform
                        const int src = __ffs(pend) - 1;
                        pend &= pend - 1;
                        const uint64_t k1 = __shfl_sync(0xffffffffu, key, src);
                        if (k1 > top[0].tau) top[0].push(k1, lane);  // re-tested: an earlier insert may have raised tau
                    }
                } else {
#pragma unroll
                for (int r = 0; r < RPW; ++r) {
                    const int64_t row = gather ? (int64_t)g_row[r] : row0 + r;
                    const bool ok = (row < p.n_rows) && ((allow_bits >> r) & 1u);
                    const float dot = warp_sum_f32(acc[0][r][0] + acc[0][r][1]);
                    if (ok) {
                        const uint64_t key = cdr_pack_key(dot * inv_q[0], (uint32_t)row);
                        if (key > top[0].tau) top[0].push(key, lane);
                    }
                }
                }
            }
        } else {
        float4 q[QPC][J];
        float inv_qn[QPC];
        WarpTopK<NPL> top[QPC];
#pragma unroll
        for (int u = 0; u < QPC; ++u) {
            float qn = 0.f;
            // a missing query of the last group re-reads the group's first query; its results are never written
            const float4 *qv = reinterpret_cast<const float4 *>(p.queries + (size_t)(qsrc0 + (u < nqv ? u : 0)) * DIM);
#pragma unroll
            for (int j = 0; j < J; ++j) {
                q[u][j] = __ldg(&qv[j * 32 + lane]);
                qn = fmaf(q[u][j].x, q[u][j].x, qn);
                qn = fmaf(q[u][j].y, q[u][j].y, qn);
                qn = fmaf(q[u][j].z, q[u][j].z, qn);
                qn = fmaf(q[u][j].w, q[u][j].w, qn);
            }
            qn = warp_sum_f32(qn);
            inv_qn[u] = __fdiv_rn(1.0f, __fsqrt_rn(qn));   // inf for a zero query -> NaN scores
            top[u].init(lists + (warp * QPC + u) * KC, lane);
        }

        for (;; ++it) {
            const int s = (int)(it % (uint32_t)S);
            const uint32_t ph = (it / (uint32_t)S) & 1u;
            mbar_wait(&full_bar[s], ph);
            const int64_t tile = stage_tile[s];
            if (tile < 0) {                                       // end of this group's stream: hand the stage back
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_bar[s]);
                ++it;
                break;
            }
            CDR_DEV_ASSERT(s >= 0 && s < S && S <= kMaxStages && tile < n_tiles);
            const int64_t row0 = tile * TR + warp * RPW;   // gather: first LIST ENTRY of this warp
            // filter bits for this warp's rows (consumed after the dot products, so the load overlaps them)
            uint32_t allow_bits = 0xFFFFFFFFu;
            uint32_t g_row[RPW];                            // gather: the rows behind the list entries
            float g_inv[RPW];
            if (gather) {
#pragma unroll
                for (int r = 0; r < RPW; ++r) {
                    const bool in = row0 + r < (int64_t)n_listed;
                    // (the 8 consumer warps overlap these two dependent L2 round trips; carrying them through the
                    // stage from the producer warp made the producer latency-bound: 0.48 vs 0.41 ms at rows/2)
                    g_row[r] = in ? __ldg(&p.row_list[row0 + r]) : 0u;
                    g_inv[r] = __ldg(&p.inv_norm[g_row[r]]);
                    if (!in) allow_bits &= ~(1u << r);
                }
            } else if (p.allow != nullptr) {
                // RPW <= 2 consecutive rows never straddle a 32-bit word (row0 is even when RPW==2)
                const uint32_t w = (row0 < p.n_rows) ? __ldg(&p.allow[row0 >> 5]) : 0u;
                allow_bits = w >> (row0 & 31);
            }

            const float4 *tv = reinterpret_cast<const float4 *>(tiles + (size_t)s * L::kTileBytes) +
                               (size_t)(warp * RPW) * (DIM / 4);
            const float *mv = reinterpret_cast<const float *>(metas + (size_t)s * L::kMetaBytes) +
                              warp * RPW;
            float acc[QPC][RPW][2];
#pragma unroll
            for (int u = 0; u < QPC; ++u)
#pragma unroll
                for (int r = 0; r < RPW; ++r) acc[u][r][0] = acc[u][r][1] = 0.f;
#pragma unroll
            for (int j = 0; j < J; ++j) {
#pragma unroll
                for (int r = 0; r < RPW; ++r) {
                    const float4 v = tv[r * (DIM / 4) + j * 32 + lane];     // one LDS.128 feeds QPC queries
#pragma unroll
                    for (int u = 0; u < QPC; ++u) {
                        float a = acc[u][r][j & 1];
                        a = fmaf(v.x, q[u][j].x, a);
                        a = fmaf(v.y, q[u][j].y, a);
                        a = fmaf(v.z, q[u][j].z, a);
                        a = fmaf(v.w, q[u][j].w, a);
                        acc[u][r][j & 1] = a;
                    }
                }
            }
            float inv_n[RPW];
#pragma unroll
            for (int r = 0; r < RPW; ++r) inv_n[r] = gather ? g_inv[r] : mv[r];
            // all reads of this stage are done (values are in registers): release it early
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[s]);

            if constexpr (QPC == 1) {
#pragma unroll
                for (int r = 0; r < RPW; ++r) {
                    const int64_t row = gather ? (int64_t)g_row[r] : row0 + r;
                    const bool ok = (row < p.n_rows) && ((allow_bits >> r) & 1u);
                    const float dot = warp_sum_f32(acc[0][r][0] + acc[0][r][1]);
                    if (ok) {
                        const float score = dot * inv_n[r] * inv_qn[0];
                        const uint64_t key = cdr_pack_key(score, (uint32_t)row);
                        if (key > top[0].tau) top[0].push(key, lane);
                    }
                }
            } else {
                // QPC*RPW (<= 8) dot products are reduced TOGETHER: at offsets 16, 8, 4 every lane keeps half of its
                // values and hands the other half to its partner, so after three exchanges the four lanes of group
                // g = lane/4 own value g, and two more butterfly steps finish it -- 9 shuffles instead of 5 per
                // value.  A lane adds exactly what the plain butterfly adds at every level (own + partner), so the
                // sums are bit-identical.  The four-lane group of value v = u*RPW + r then scores, packs and tests
                // its candidate in parallel; the rare survivors are inserted one by one.
                constexpr int V = QPC * RPW;
                static_assert(V <= 8, "the transposed reduction handles at most 8 values");
                float a8[8];
#pragma unroll
                for (int v = 0; v < 8; ++v) a8[v] = v < V ? acc[v / RPW][v % RPW][0] + acc[v / RPW][v % RPW][1] : 0.f;
                const bool b4 = (lane & 16) != 0, b3 = (lane & 8) != 0, b2 = (lane & 4) != 0;
                float h4[4], h2[2];
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    h4[i] = (b4 ? a8[i + 4] : a8[i]) + __shfl_xor_sync(0xffffffffu, b4 ? a8[i] : a8[i + 4], 16);
#pragma unroll
                for (int i = 0; i < 2; ++i)
                    h2[i] = (b3 ? h4[i + 2] : h4[i]) + __shfl_xor_sync(0xffffffffu, b3 ? h4[i] : h4[i + 2], 8);
                float dot = (b2 ? h2[1] : h2[0]) + __shfl_xor_sync(0xffffffffu, b2 ? h2[0] : h2[1], 4);
                dot += __shfl_xor_sync(0xffffffffu, dot, 2);
                dot += __shfl_xor_sync(0xffffffffu, dot, 1);
                const int v = (lane >> 2) & 7;                 // the value this lane group owns
                const int vu = v / RPW, vr = v % RPW;
                float s_inv_n = inv_n[0], s_inv_q = inv_qn[0];
                uint64_t s_tau = top[0].tau;
                int64_t s_row = gather ? (int64_t)g_row[0] : row0;
#pragma unroll
                for (int r = 1; r < RPW; ++r)
                    if (vr == r) { s_inv_n = inv_n[r]; s_row = gather ? (int64_t)g_row[r] : row0 + r; }
#pragma unroll
                for (int u = 1; u < QPC; ++u)
                    if (vu == u) { s_inv_q = inv_qn[u]; s_tau = top[u].tau; }
                const bool s_ok = v < V && (s_row < p.n_rows) && ((allow_bits >> vr) & 1u);
                const uint64_t key = cdr_pack_key(dot * s_inv_n * s_inv_q, (uint32_t)s_row);
                unsigned pend = __ballot_sync(0xffffffffu, s_ok && key > s_tau) & 0x11111111u;
                while (pend) {                                 // rare; warp-uniform
                    const int src = __ffs(pend) - 1;
                    pend &= pend - 1;
                    const uint64_t k1 = __shfl_sync(0xffffffffu, key, src);
                    const int su = (src >> 2) / RPW;
#pragma unroll
                    for (int u = 0; u < QPC; ++u)
                        if (su == u && k1 > top[u].tau) top[u].push(k1, lane);   // re-tested: an earlier insert may have raised tau
                }
            }
        }

        }

        // ---- per query: per-warp sort, then 3-level pairwise merge through shared memory
#pragma unroll 1
        for (int u = 0; u < (L::kDeep ? 0 : QPC); ++u) {
            uint64_t *mine = lists + (warp * QPC + u) * KC;
            uint64_t k[NPL];
            __syncwarp();
#pragma unroll
            for (int i = 0; i < NPL; ++i) k[i] = mine[i * 32 + lane];
            warp_bitonic_sort_desc<NPL>(k, lane);
#pragma unroll
            for (int i = 0; i < NPL; ++i) mine[i * 32 + lane] = k[i];
#pragma unroll
            for (int step = 1; step < CW; step <<= 1) {
                asm volatile("bar.sync 1, %0;" ::"n"(CW * 32) : "memory");
                if ((warp & (2 * step - 1)) == 0) {
                    warp_merge_topk<NPL>(k, lists + ((warp + step) * QPC + u) * KC, lane);
                }
                asm volatile("bar.sync 1, %0;" ::"n"(CW * 32) : "memory");
                if ((warp & (2 * step - 1)) == 0) {
#pragma unroll
                    for (int i = 0; i < NPL; ++i) mine[i * 32 + lane] = k[i];
                }
            }
            if (warp == 0 && u < nqv) {
                CDR_DEV_ASSERT(q0 + u < nq_eff && p.list_offset + (int)blockIdx.x < p.lists_per_query);
                uint64_t *out = p.cta_keys + ((size_t)(q0 + u) * p.lists_per_query + p.list_offset + blockIdx.x) * KC;
#pragma unroll
                for (int i = 0; i < NPL; ++i) out[i * 32 + lane] = k[i];
            }
        }
        // the next group re-initialises the lists (and, deep reads, the staged queries): every merge that reads
        // another warp's list ended before the last barrier above
        }
    }
}

// Allowed rows of a filter bitmap as a compact list (any order: candidate keys carry the row, so the
// result does not depend on it) + their exact count.  One warp per 32 words; one atomic per warp-iteration.
__global__ void __launch_bounds__(256) compact_allow_kernel(const uint32_t *allow, int64_t n_rows, uint32_t *list,
                                                            uint32_t cap, unsigned int *count)
{
    const int lane = threadIdx.x & 31;
    const int64_t words = (n_rows + 31) >> 5;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t w0 = warp0 * 32; w0 < words; w0 += nwarps * 32) {
        const int64_t w = w0 + lane;
        uint32_t bits = w < words ? allow[w] : 0u;
        if (w == words - 1 && (n_rows & 31)) bits &= (1u << (n_rows & 31)) - 1u;   // caller bitmaps may carry tail bits
        const int c = __popc(bits);
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        if (total == 0) continue;
        unsigned int base = 0;
        if (lane == 0) base = atomicAdd(count, (unsigned int)total);
        base = __shfl_sync(0xffffffffu, base, 0);
        unsigned int pos = base + (unsigned int)(incl - c);
        while (bits) {
            const int b = __ffs(bits) - 1;
            bits &= bits - 1;
            if (pos < cap) list[pos] = (uint32_t)((w << 5) + b);
            ++pos;
        }
    }
}

// ------------------------------------------------------------------------------------------
// Finalize: one CTA (8 warps) per query.
//   1. k-way merge: warp w folds lists w, w+8, ... (each sorted desc, KC keys) into its registers
//      with the bitonic top-k merge, then the 8 partial results are merged pairwise.
//   2. fp64 re-score of the KC survivors on the resident fp32 rows (pgvector's formula with
//      fp64 accumulators): sim = ab / sqrt(aa*bb), clamp, distance = 1 - sim, score = 1 - distance.
//   3. final order (score desc, NaN last, id asc) by rank counting; first k written out.
struct FinalizeParams {
    const uint64_t *lists;    // sorted mode: [nq, n_lists, KC]; unsorted mode: [nq, cap]
    int n_lists;
    const uint32_t *counts;   // unsorted mode: valid keys per query (clamped to cap); else nullptr
    int cap;
    const float *rows;        // fp32 rows (may be null when bf16_rows is used)
    const __nv_bfloat16 *bf16_rows;
    const float *queries;     // [nq, dim]
    const int64_t *ids;
    int dim;
    int k;
    double *out_score;        // [nq, k]
    int64_t *out_id;          // [nq, k]
    int32_t *out_n;           // [nq]
    unsigned int *reset_ctr;  // nullable: K1's work-stealing counter of this query's group, zeroed for the next launch
    int reset_div;            // queries per K1 group (counter index = query / reset_div)
    // conditional re-run (see ScanParams): lists of SLOT g belong to query q_index[g]; min(*q_count, n_slots) slots are
    // live and the CTAs of the launch walk them; both null = slot == query, one CTA per query
    const int *q_index;
    const int *q_count;
    int n_slots;
    // PEER instantiations (plain launches of a row-sharded step): the CTA that orders a query's list goes on to exchange it
    // with the other ranks (peer_exchange_cta) and out_* receive the MERGED lists -- the K4p kernel folded into this one
    PeerLink peer;
};

// Shared memory of the fused exchange: this rank's ordered list and the merge scratch (world <= 8 ranks x k <= 64).
template <bool PEER>
struct PeerFusedSmem {
    double l_sc[PEER ? kPeerFusedK : 1];
    int64_t l_id[PEER ? kPeerFusedK : 1];
    uint64_t m_key[PEER ? kPeerFusedRanks * kPeerFusedK : 1];
    int64_t m_id[PEER ? kPeerFusedRanks * kPeerFusedK : 1];
    int n[kPeerFusedRanks];
    int cnt;
};

// fp64 cosine of query (a-values in smem as float) x one resident row; all lanes return the result.
__device__ __forceinline__ void rescore_accumulate(const float4 a, const float4 b, double &ab, double &bb)
{
    ab = __fma_rn((double)a.x, (double)b.x, ab);
    ab = __fma_rn((double)a.y, (double)b.y, ab);
    ab = __fma_rn((double)a.z, (double)b.z, ab);
    ab = __fma_rn((double)a.w, (double)b.w, ab);
    bb = __fma_rn((double)b.x, (double)b.x, bb);
    bb = __fma_rn((double)b.y, (double)b.y, bb);
    bb = __fma_rn((double)b.z, (double)b.z, bb);
    bb = __fma_rn((double)b.w, (double)b.w, bb);
}
__device__ __forceinline__ float4 load_row_vec(const float *rows, const __nv_bfloat16 *bf16_rows, size_t row,
                                               int dim, int v)
{
    if (rows != nullptr) return __ldg(reinterpret_cast<const float4 *>(rows + row * dim) + v);
    const uint2 raw = __ldg(reinterpret_cast<const uint2 *>(bf16_rows + row * dim) + v);
    return make_float4(__uint_as_float(raw.x << 16), __uint_as_float(raw.x & 0xFFFF0000u),
                       __uint_as_float(raw.y << 16), __uint_as_float(raw.y & 0xFFFF0000u));
}

template <int NPL, int WARPS, int kRB, bool PEER = false>
__global__ void __launch_bounds__(WARPS * 32, WARPS <= 8 ? 4 : 1) scan_finalize_kernel(const FinalizeParams p)
{
    constexpr int KC = NPL * 32;
    __shared__ uint64_t s_lists[WARPS * KC];
    __shared__ double s_score[KC];
    __shared__ uint64_t s_key[KC];            // order key of s_score (0 for an empty slot): the ranking compares integers
    __shared__ int64_t s_id[KC];
    __shared__ int s_valid[KC];
    __shared__ int s_nvalid;
    __shared__ PeerFusedSmem<PEER> s_peer;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    asm volatile("griddepcontrol.wait;" ::: "memory");      // launched beside the scan (see exact_scan_kernel): its lists are complete now
    int n_live = p.n_slots;
    if (p.q_count != nullptr) {
        const int c = *p.q_count;
        n_live = c < n_live ? c : n_live;
    }
    for (int slot = blockIdx.x; slot < n_live; slot += gridDim.x) {
    const int qi = p.q_index != nullptr ? __ldg(&p.q_index[slot]) : slot;     // query: inputs and outputs
    CDR_DEV_ASSERT(qi >= 0 && slot < p.n_slots);
    const uint64_t *base = p.lists + (size_t)slot * p.n_lists * KC;
    __syncthreads();                                                          // the previous slot's shared state is dead
    if (threadIdx.x == 0) s_nvalid = 0;

    uint64_t k[NPL];
#pragma unroll
    for (int i = 0; i < NPL; ++i) k[i] = CDR_EMPTY_KEY;
    if (p.counts == nullptr) {
        // sorted lists: software-pipelined so the next list's keys are in flight during a merge
        uint64_t nxt[NPL];
        int l = warp;
        if (l < p.n_lists) {
#pragma unroll
            for (int i = 0; i < NPL; ++i) nxt[i] = base[(size_t)l * KC + (KC - 1 - (i * 32 + lane))];
        }
        while (l < p.n_lists) {
            uint64_t cur[NPL];
#pragma unroll
            for (int i = 0; i < NPL; ++i) cur[i] = nxt[i];
            const int ln = l + WARPS;
            if (ln < p.n_lists) {
#pragma unroll
                for (int i = 0; i < NPL; ++i) nxt[i] = base[(size_t)ln * KC + (KC - 1 - (i * 32 + lane))];
            }
#pragma unroll
            for (int i = 0; i < NPL; ++i) k[i] = k[i] > cur[i] ? k[i] : cur[i];   // cur is already reversed
            warp_bitonic_merge_desc<NPL>(k, lane);
            l = ln;
        }
    } else {
        // one unsorted list per query (K2 candidates): sort KC-sized chunks, fold them in
        const uint64_t *list = p.lists + (size_t)slot * p.cap;
        uint32_t n = p.counts[slot];
        if (n > (uint32_t)p.cap) n = p.cap;
        const int chunks = (int)((n + KC - 1) / KC);
        for (int ch = warp; ch < chunks; ch += WARPS) {
            uint64_t c[NPL];
#pragma unroll
            for (int i = 0; i < NPL; ++i) {
                const uint32_t e = (uint32_t)ch * KC + i * 32 + lane;
                c[i] = e < n ? list[e] : CDR_EMPTY_KEY;
            }
            warp_bitonic_sort_desc<NPL>(c, lane);
#pragma unroll
            for (int i = 0; i < NPL; ++i) s_lists[warp * KC + i * 32 + lane] = c[i];
            __syncwarp();
            warp_merge_topk<NPL>(k, s_lists + warp * KC, lane);
            __syncwarp();
        }
    }
#pragma unroll
    for (int i = 0; i < NPL; ++i) s_lists[warp * KC + i * 32 + lane] = k[i];
#pragma unroll
    for (int step = 1; step < WARPS; step <<= 1) {
        __syncthreads();
        if ((warp & (2 * step - 1)) == 0) {
            warp_merge_topk<NPL>(k, s_lists + (warp + step) * KC, lane);
#pragma unroll
            for (int i = 0; i < NPL; ++i) s_lists[warp * KC + i * 32 + lane] = k[i];   // own slot: nobody reads it at this level
        }
    }
    __syncthreads();

    // ---- fp64 re-score: warp w takes candidates w, w+WARPS, ...; two rows in flight per warp
    const float *qrow = p.queries + (size_t)qi * p.dim;
    const int nvec = p.dim >> 2;   // float4 per row
    // Loads are issued in batches of kRB independent 16-byte vectors per lane (and per row below), so a
    // candidate row costs one or two memory round trips instead of nvec/32 serialised ones; the per-lane
    // accumulation order (v = lane, lane+32, ...) -- and therefore every bit of the result -- is unchanged.
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    double aa = 0.0;
    for (int v0 = lane; v0 < nvec; v0 += 32 * 2 * kRB) {
        float4 a[2 * kRB];
#pragma unroll
        for (int u = 0; u < 2 * kRB; ++u) {
            const int v = v0 + 32 * u;
            a[u] = v < nvec ? __ldg(reinterpret_cast<const float4 *>(qrow) + v) : zero4;
        }
#pragma unroll
        for (int u = 0; u < 2 * kRB; ++u) {
            aa = __fma_rn((double)a[u].x, (double)a[u].x, aa);
            aa = __fma_rn((double)a[u].y, (double)a[u].y, aa);
            aa = __fma_rn((double)a[u].z, (double)a[u].z, aa);
            aa = __fma_rn((double)a[u].w, (double)a[u].w, aa);
        }
    }
    aa = warp_sum_f64(aa);
    int my_valid = 0;
    for (int c0 = warp; c0 < KC; c0 += 2 * WARPS) {
        const int c1 = c0 + WARPS;
        const uint64_t key0 = s_lists[c0];
        const uint64_t key1 = c1 < KC ? s_lists[c1] : CDR_EMPTY_KEY;
        const bool ok0 = key0 != CDR_EMPTY_KEY, ok1 = key1 != CDR_EMPTY_KEY;
        const size_t row0 = ok0 ? cdr_key_row(key0) : 0, row1 = ok1 ? cdr_key_row(key1) : 0;
        double ab0 = 0.0, bb0 = 0.0, ab1 = 0.0, bb1 = 0.0;
        if (ok0 | ok1) {
            for (int v0 = lane; v0 < nvec; v0 += 32 * kRB) {
                float4 a[kRB], b0[kRB], b1[kRB];
#pragma unroll
                for (int u = 0; u < kRB; ++u) {
                    const int v = v0 + 32 * u;
                    const bool in = v < nvec;
                    b0[u] = in ? load_row_vec(p.rows, p.bf16_rows, row0, p.dim, v) : zero4;
                    b1[u] = in ? load_row_vec(p.rows, p.bf16_rows, row1, p.dim, v) : zero4;
                    a[u] = in ? __ldg(reinterpret_cast<const float4 *>(qrow) + v) : zero4;
                }
#pragma unroll
                for (int u = 0; u < kRB; ++u) {
                    rescore_accumulate(a[u], b0[u], ab0, bb0);
                    rescore_accumulate(a[u], b1[u], ab1, bb1);
                }
            }
        }
        ab0 = warp_sum_f64(ab0); bb0 = warp_sum_f64(bb0);
        ab1 = warp_sum_f64(ab1); bb1 = warp_sum_f64(bb1);
        if (lane < 2) {
            const int c = lane == 0 ? c0 : c1;
            const bool ok = lane == 0 ? ok0 : ok1;
            if (c < KC) {
                if (ok) {
                    const double ab = lane == 0 ? ab0 : ab1, bb = lane == 0 ? bb0 : bb1;
                    double sim = __ddiv_rn(ab, __dsqrt_rn(__dmul_rn(aa, bb)));
                    if (sim > 1.0) sim = 1.0;
                    else if (sim < -1.0) sim = -1.0;
                    const double dist = __dsub_rn(1.0, sim);       // pgvector cosine_distance (float8)
                    const double score = __dsub_rn(1.0, dist);     // SQL: 1 - (embedding <=> q)
                    s_score[c] = score;
                    s_key[c] = cdr_order_f64(score);
                    s_id[c] = p.ids[lane == 0 ? row0 : row1];
                    s_valid[c] = 1;
                    my_valid += 1;
                } else {
                    s_valid[c] = 0; s_score[c] = 0.0; s_key[c] = 0ull; s_id[c] = INT64_MAX;
                }
            }
        }
    }
    if (my_valid) atomicAdd(&s_nvalid, my_valid);
    __syncthreads();

    // ---- final order by rank counting: KC candidates x PARTS threads each (a thread counts every PARTS-th rival)
    const int n_valid = s_nvalid;
    constexpr int PARTS = (WARPS * 32) / KC >= 32 ? 32 : ((WARPS * 32) / KC >= 1 ? (WARPS * 32) / KC : 1);
    static_assert((PARTS & (PARTS - 1)) == 0 && KC * PARTS <= WARPS * 32, "rank-count split");
    for (int t = threadIdx.x; t < KC * PARTS; t += blockDim.x) {
        const int c = t / PARTS, part = t % PARTS;
        const bool live = s_valid[c] != 0;
        const uint64_t key = s_key[c];
        const int64_t id = s_id[c];
        // branch-free: an empty slot carries (key 0, id INT64_MAX) and is never "before" anything, nor is c before itself
        int rank = 0;
#pragma unroll 16
        for (int j = 0; j < KC / PARTS; ++j) {
            const int o = part + j * PARTS;
            const uint64_t ko = s_key[o];
            const int64_t io = s_id[o];
            rank += (int)((ko > key) | ((ko == key) & (io < id)));
        }
#pragma unroll
        for (int d = PARTS >> 1; d > 0; d >>= 1) rank += __shfl_xor_sync(0xffffffffu, rank, d);
        if (!live || part != 0) continue;
        const double sc = s_score[c];
        if (rank < p.k) {
            if constexpr (PEER) {
                s_peer.l_sc[rank] = sc;
                s_peer.l_id[rank] = id;
            } else {
                p.out_score[(size_t)qi * p.k + rank] = sc;
                p.out_id[(size_t)qi * p.k + rank] = id;
            }
        }
    }
    const int n_out = n_valid < p.k ? n_valid : p.k;
    if constexpr (PEER) {
        if (threadIdx.x == 0 && p.reset_ctr) p.reset_ctr[slot / p.reset_div] = 0u;
        peer_exchange_cta(p.peer, p.peer.q0 + qi, p.k, s_peer.l_sc, s_peer.l_id, n_out, s_peer.m_key, s_peer.m_id, s_peer.n,
                          &s_peer.cnt, p.out_score + (size_t)qi * p.k, p.out_id + (size_t)qi * p.k, p.out_n + qi);
    } else {
    for (int c = n_out + threadIdx.x; c < p.k; c += blockDim.x) {
        p.out_score[(size_t)qi * p.k + c] = __longlong_as_double(0x7FF8000000000000ll);
        p.out_id[(size_t)qi * p.k + c] = -1;
    }
    if (threadIdx.x == 0) {
        p.out_n[qi] = n_out;
        if (p.reset_ctr) p.reset_ctr[slot / p.reset_div] = 0u;
    }
    }
    }
}

// ------------------------------------------------------------------------------------------
// Latency variant of the finalize step for small batches (one request = one query): a cluster of
// 8 CTAs per query.  A single CTA is issue-bound (~60 k warp-instructions on one SM, 25 us); here
//   1. every CTA merges its share of the per-CTA lists (lists r, r+8, ...) to a top-KC in shared memory,
//   2. CTA 0 reads the 8 partial lists through distributed shared memory and merges them,
//   3. every CTA re-scores KC/8 of the survivors in fp64 (one row per warp) and stores score / id into
//      CTA 0's shared memory,
//   4. CTA 0 orders the KC results by rank counting and writes the first k.
// Same arithmetic and ordering as scan_finalize_kernel => identical bits.
constexpr int kFinCluster = 8;

// -DCDR_FIN_TIMING: thread 0 of the cluster's first CTA prints the kernel's phase boundaries (globaltimer, ns) -- a probe
// build (csrc/build_ab.sh), compiled out of the shipped library.
#ifdef CDR_FIN_TIMING
#define CDR_FIN_STAMP(i) do { if (threadIdx.x == 0 && blockIdx.x == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); fin_t[i] = t_; } } while (0)
#else
#define CDR_FIN_STAMP(i) do { } while (0)
#endif

// U: 16-byte vectors of a row per lane (dim <= 128 U); the U = 8 instantiation (dim <= 1024) fits beside a scan CTA (36 K of
// the SM's 64 K registers), so the programmatically launched finalize is resident and through its prologue while the scan runs.
template <int NPL, bool PEER = false, int U = 16>
__global__ void __cluster_dims__(kFinCluster, 1, 1) __launch_bounds__(256) __maxnreg__(U == 8 ? 112 : 255)
    scan_finalize_cluster_kernel(const FinalizeParams p)
{
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    constexpr int KC = NPL * 32, WARPS = 8;
    __shared__ uint64_t s_lists[WARPS * KC];   // phase 1: this CTA's partial lists; [0, KC) = its top-KC
    __shared__ uint64_t s_tree[WARPS * KC];    // phase 2 (CTA 0): the 8 CTAs' lists, pushed by their owners
    __shared__ uint64_t s_surv[KC];            // phase 3: the survivors this CTA re-scores (slots c of its warps), pushed by CTA 0
    __shared__ double s_score[KC];             // phase 3 results, written into CTA 0 by all CTAs
    __shared__ uint64_t s_key[KC];             // order key of s_score (0 for an empty slot): the ranking compares integers
    __shared__ int64_t s_id[KC];
    __shared__ int s_nvalid;
    __shared__ __align__(8) uint64_t s_bar[3];  // arrival of: the 8 lists (CTA 0), my survivors, the KC results (CTA 0)
    __shared__ PeerFusedSmem<PEER> s_peer;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned crank = cluster.block_rank();
#ifdef CDR_FIN_TIMING
    unsigned long long fin_t[16] = {0};
#endif
    CDR_FIN_STAMP(0);
    const int qi = blockIdx.x / kFinCluster;
    const uint64_t *base = p.lists + (size_t)qi * p.n_lists * KC;
    if (threadIdx.x == 0) s_nvalid = 0;

    // The query row (registers, 16-byte vectors v = lane + 32 u) and its fp64 norm need nothing from the scan: both are
    // ready before the dependency wait.
    const float *qrow = p.queries + (size_t)qi * p.dim;
    const int nvec = p.dim >> 2;
    constexpr int kMaxU = U;                       // dim <= 128 U
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 a[kMaxU];
#pragma unroll
    for (int u = 0; u < kMaxU; ++u) {
        const int v = lane + 32 * u;
        a[u] = v < nvec ? __ldg(reinterpret_cast<const float4 *>(qrow) + v) : zero4;
    }
    double aa = 0.0;
#pragma unroll
    for (int u = 0; u < kMaxU; ++u) {
        if (32 * u < nvec) {                       // (vectors past the row end are zero: they would add nothing)
            aa = __fma_rn((double)a[u].x, (double)a[u].x, aa);
            aa = __fma_rn((double)a[u].y, (double)a[u].y, aa);
            aa = __fma_rn((double)a[u].z, (double)a[u].z, aa);
            aa = __fma_rn((double)a[u].w, (double)a[u].w, aa);
        }
    }
    aa = warp_sum_f64(aa);
    // Every exchange between the CTAs of the cluster is a PUSH: st.async stores into the receiver's shared memory that
    // complete bytes on an mbarrier there, armed here (a cluster barrier costs 2.5-4 us on this path, profiles/r02/latency;
    // the one below, which publishes the barriers, runs before the dependency wait).  Each barrier serves one phase.
    if (threadIdx.x == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        mbar_init(&s_bar[2], 1);
        fence_mbar_init();
        if (crank == 0) mbar_arrive_expect_tx(&s_bar[0], (uint32_t)(kFinCluster * KC * 8));
        mbar_arrive_expect_tx(&s_bar[1], (uint32_t)(KC / kFinCluster * 8));
        if (crank == 0) mbar_arrive_expect_tx(&s_bar[2], (uint32_t)(KC * 24));
    }
    cluster.sync();
    const uint32_t bar0_tree = mapa_u32(smem_u32(&s_bar[0]), 0), bar0_res = mapa_u32(smem_u32(&s_bar[2]), 0);
    asm volatile("griddepcontrol.wait;" ::: "memory");      // launched beside the scan (see exact_scan_kernel): its lists are complete now
    CDR_FIN_STAMP(1);

    // ---- 1. this CTA's lists: l = crank + 8 * (warp + 8 * j), software-pipelined like the 1-CTA kernel
    uint64_t k[NPL];
#pragma unroll
    for (int i = 0; i < NPL; ++i) k[i] = CDR_EMPTY_KEY;
    {
        constexpr int STRIDE = kFinCluster * WARPS;
        uint64_t nxt[NPL];
        int l = (int)crank + kFinCluster * warp;
        if (l < p.n_lists) {
#pragma unroll
            for (int i = 0; i < NPL; ++i) nxt[i] = base[(size_t)l * KC + (KC - 1 - (i * 32 + lane))];
        }
        while (l < p.n_lists) {
            uint64_t cur[NPL];
#pragma unroll
            for (int i = 0; i < NPL; ++i) cur[i] = nxt[i];
            const int ln = l + STRIDE;
            if (ln < p.n_lists) {
#pragma unroll
                for (int i = 0; i < NPL; ++i) nxt[i] = base[(size_t)ln * KC + (KC - 1 - (i * 32 + lane))];
            }
#pragma unroll
            for (int i = 0; i < NPL; ++i) k[i] = k[i] > cur[i] ? k[i] : cur[i];
            warp_bitonic_merge_desc<NPL>(k, lane);
            l = ln;
        }
    }
#pragma unroll
    for (int i = 0; i < NPL; ++i) s_lists[warp * KC + i * 32 + lane] = k[i];
#pragma unroll
    for (int step = 1; step < WARPS; step <<= 1) {
        __syncthreads();
        if ((warp & (2 * step - 1)) == 0) {
            warp_merge_topk<NPL>(k, s_lists + (warp + step) * KC, lane);
            if (step * 2 < WARPS) {
#pragma unroll
                for (int i = 0; i < NPL; ++i) s_lists[warp * KC + i * 32 + lane] = k[i];
            }
        }
    }
    if (warp == 0) {                              // this CTA's top-KC goes to CTA 0 as list `crank` of its tree
#pragma unroll
        for (int i = 0; i < NPL; ++i)
            st_async_b64(mapa_u32(smem_u32(&s_tree[crank * KC + i * 32 + lane]), 0), k[i], bar0_tree);
    }
    CDR_FIN_STAMP(2);

    // ---- 2. CTA 0: the same pairwise tree over the 8 CTAs' lists; warp 0 ends with the KC survivors, sorted, and hands
    //         survivor c to its owner CTA (c mod 64) / 8, which re-scores it in warp c mod 8
    if (crank == 0) {
        mbar_wait(&s_bar[0], 0);
#pragma unroll
        for (int i = 0; i < NPL; ++i) k[i] = s_tree[warp * KC + i * 32 + lane];
#pragma unroll
        for (int step = 1; step < WARPS; step <<= 1) {
            if (step > 1) __syncthreads();
            if ((warp & (2 * step - 1)) == 0) {
                warp_merge_topk<NPL>(k, s_tree + (warp + step) * KC, lane);
                if (step * 2 < WARPS) {
#pragma unroll
                    for (int i = 0; i < NPL; ++i) s_tree[warp * KC + i * 32 + lane] = k[i];
                }
            }
        }
        if (warp == 0) {
#pragma unroll
            for (int i = 0; i < NPL; ++i) {
                const int c = i * 32 + lane;
                const uint32_t owner = (uint32_t)((c % (kFinCluster * WARPS)) / WARPS);
                st_async_b64(mapa_u32(smem_u32(&s_surv[c]), owner), k[i], mapa_u32(smem_u32(&s_bar[1]), owner));
            }
        }
    }
    mbar_wait(&s_bar[1], 0);
    CDR_FIN_STAMP(3);

    // ---- 3. fp64 re-score, one survivor per warp at a time: c = crank*8 + warp + 64*j; the row is read in ONE batch of
    //         independent 16-byte loads per lane (the query vectors are in registers), accumulated in the order
    //         v = lane, lane + 32, ... of scan_finalize_kernel => the same bits
    CDR_FIN_STAMP(15);
    for (int c = (int)crank * WARPS + warp; c < KC; c += kFinCluster * WARPS) {
        const uint64_t key = s_surv[c];
        const bool ok = key != CDR_EMPTY_KEY;
        const size_t row = ok ? cdr_key_row(key) : 0;
        CDR_FIN_STAMP(7);
        int64_t row_id = INT64_MAX;
        if (ok && lane == 0) row_id = __ldg(&p.ids[row]);
        float4 b[kMaxU];
#pragma unroll
        for (int u = 0; u < kMaxU; ++u) {
            const int v = lane + 32 * u;
            b[u] = (ok && v < nvec) ? load_row_vec(p.rows, p.bf16_rows, row, p.dim, v) : zero4;
            // (the widest candidate lists leave no registers to hold the query across the merges: read again, L1/L2 hits)
            if constexpr (NPL > 4) a[u] = (ok && v < nvec) ? __ldg(reinterpret_cast<const float4 *>(qrow) + v) : zero4;
        }
        double ab = 0.0, bb = 0.0;
#pragma unroll
        for (int u = 0; u < kMaxU; ++u) {
            if (ok && 32 * u < nvec) rescore_accumulate(a[u], b[u], ab, bb);
        }
        CDR_FIN_STAMP(8);
        ab = warp_sum_f64(ab);
        bb = warp_sum_f64(bb);
        CDR_FIN_STAMP(9);
        if (lane == 0) {                          // result c -> CTA 0: score, its order key, id (INT64_MAX = empty slot)
            double score = 0.0;
            if (ok) {
                double sim = __ddiv_rn(ab, __dsqrt_rn(__dmul_rn(aa, bb)));
                if (sim > 1.0) sim = 1.0;
                else if (sim < -1.0) sim = -1.0;
                const double dist = __dsub_rn(1.0, sim);
                score = __dsub_rn(1.0, dist);
            }
            CDR_FIN_STAMP(10);
            st_async_b64(mapa_u32(smem_u32(&s_score[c]), 0), (uint64_t)__double_as_longlong(score), bar0_res);
            st_async_b64(mapa_u32(smem_u32(&s_key[c]), 0), ok ? cdr_order_f64(score) : 0ull, bar0_res);
            st_async_b64(mapa_u32(smem_u32(&s_id[c]), 0), (uint64_t)row_id, bar0_res);
        }
    }
    CDR_FIN_STAMP(11);
    if (crank != 0) return;                       // (its pushes are posted; nobody reads this CTA's memory any more)
    mbar_wait(&s_bar[2], 0);
    CDR_FIN_STAMP(4);

    // ---- 4. CTA 0: final order by rank counting
    int mine = 0;
    for (int c = threadIdx.x; c < KC; c += blockDim.x) mine += s_id[c] != INT64_MAX;
    if (mine) atomicAdd(&s_nvalid, mine);
    CDR_FIN_STAMP(12);
    __syncthreads();
    CDR_FIN_STAMP(13);
    const int n_valid = s_nvalid;
    constexpr int PARTS = 256 / KC >= 1 ? 256 / KC : 1;       // threads per candidate: each counts every PARTS-th rival
    for (int t = threadIdx.x; t < KC * PARTS; t += blockDim.x) {
        const int c = t / PARTS, part = t % PARTS;
        const bool live = s_id[c] != INT64_MAX;
        const uint64_t key = s_key[c];
        const int64_t id = s_id[c];
        // branch-free: an empty slot carries (key 0, id INT64_MAX) and is never "before" anything, nor is c before itself
        int rank = 0;
#pragma unroll 16
        for (int j = 0; j < KC / PARTS; ++j) {
            const int o = part + j * PARTS;
            const uint64_t ko = s_key[o];
            const int64_t io = s_id[o];
            rank += (int)((ko > key) | ((ko == key) & (io < id)));
        }
        CDR_FIN_STAMP(14);
#pragma unroll
        for (int d = PARTS >> 1; d > 0; d >>= 1) rank += __shfl_xor_sync(0xffffffffu, rank, d);
        if (!live || part != 0) continue;
        const double sc = s_score[c];
        if (rank < p.k) {
            if constexpr (PEER) {
                s_peer.l_sc[rank] = sc;
                s_peer.l_id[rank] = id;
            } else {
                p.out_score[(size_t)qi * p.k + rank] = sc;
                p.out_id[(size_t)qi * p.k + rank] = id;
            }
        }
    }
    const int n_out = n_valid < p.k ? n_valid : p.k;
    CDR_FIN_STAMP(5);
    if constexpr (PEER) {
        if (threadIdx.x == 0 && p.reset_ctr) p.reset_ctr[qi / p.reset_div] = 0u;
        peer_exchange_cta(p.peer, p.peer.q0 + qi, p.k, s_peer.l_sc, s_peer.l_id, n_out, s_peer.m_key, s_peer.m_id, s_peer.n,
                          &s_peer.cnt, p.out_score + (size_t)qi * p.k, p.out_id + (size_t)qi * p.k, p.out_n + qi);
    } else {
        for (int c = n_out + threadIdx.x; c < p.k; c += blockDim.x) {
            p.out_score[(size_t)qi * p.k + c] = __longlong_as_double(0x7FF8000000000000ll);
            p.out_id[(size_t)qi * p.k + c] = -1;
        }
        if (threadIdx.x == 0) {
            p.out_n[qi] = n_out;
            if (p.reset_ctr) p.reset_ctr[qi / p.reset_div] = 0u;
        }
    }
    CDR_FIN_STAMP(6);
#ifdef CDR_FIN_TIMING
    if (threadIdx.x == 0 && blockIdx.x == 0)
        printf("FIN start->wait %llu | lists+tree %llu | cta0 tree %llu | rescore %llu (map %llu, key %llu, rows+fma %llu, sums %llu, div %llu, store+id %llu, "
               "sync %llu) | rank %llu (count %llu, bar %llu, loop %llu, shfl+store %llu) | out/exchange %llu | since-release %llu ns\n",
               fin_t[1] - fin_t[0], fin_t[2] - fin_t[1], fin_t[3] - fin_t[2], fin_t[4] - fin_t[3], fin_t[15] - fin_t[3], fin_t[7] - fin_t[15],
               fin_t[8] - fin_t[7], fin_t[9] - fin_t[8], fin_t[10] - fin_t[9], fin_t[11] - fin_t[10], fin_t[4] - fin_t[11],
               fin_t[5] - fin_t[4], fin_t[12] - fin_t[4], fin_t[13] - fin_t[12], fin_t[14] - fin_t[13], fin_t[5] - fin_t[14],
               fin_t[6] - fin_t[5], fin_t[6] - fin_t[1]);
#endif
}

// pdl: the launch may be scheduled while the kernel before it on `st` (the scan) is still running -- the finalize kernels
// wait for that grid themselves (griddepcontrol.wait) before they read its lists.
template <typename Kern>
static void launch_fin(Kern kern, unsigned grid, unsigned block, cudaStream_t st, bool pdl, const FinalizeParams &fp)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid, 1, 1);
    cfg.blockDim = dim3(block, 1, 1);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kern, fp);        // the caller reads the launch status (CDR_LAUNCH_CHECK)
}

// one launch helper for every candidate width
static int launch_finalize(const FinalizeParams &fp, int kc, int nq, cudaStream_t st, bool pdl = false)
{
    // load batch kRB: 2 fits the 64-register budget of the 1024-thread variants (4 spills; measured
    // 25.5 us vs 31.6 us per launch, profiles/r01/README.md)
    // small batches of sorted per-CTA lists (the exact lane serving single requests): 8-CTA cluster per query
    static const bool no_cluster = [] { const char *e = getenv("CADENCE_FIN_CLUSTER"); return e && e[0] == '0'; }();
    if (fp.peer.world > 0) {
        // fused exchange (cdr_scan_peer_fusable checked the shape): sorted per-CTA lists, KC 64 or 128
        if (fp.counts != nullptr || fp.q_index != nullptr || (kc != 64 && kc != 128)) {
            cdr_set_error("finalize: fused exchange not built for this launch shape (kc %d)", kc);
            return CDR_ERR_UNSUPPORTED;
        }
        if (nq <= 16 && !no_cluster) {
            if (kc == 64 && fp.dim <= 1024) launch_fin(scan_finalize_cluster_kernel<2, true, 8>, nq * kFinCluster, 256, st, pdl, fp);
            else if (kc == 64) launch_fin(scan_finalize_cluster_kernel<2, true>, nq * kFinCluster, 256, st, pdl, fp);
            else if (fp.dim <= 1024) launch_fin(scan_finalize_cluster_kernel<4, true, 8>, nq * kFinCluster, 256, st, pdl, fp);
            else launch_fin(scan_finalize_cluster_kernel<4, true>, nq * kFinCluster, 256, st, pdl, fp);
        } else {
            if (kc == 64) launch_fin(scan_finalize_kernel<2, 32, 2, true>, nq, 1024, st, pdl, fp);
            else launch_fin(scan_finalize_kernel<4, 32, 2, true>, nq, 1024, st, pdl, fp);
        }
        CDR_LAUNCH_CHECK();
        return CDR_OK;
    }
    if (fp.counts == nullptr && fp.q_index == nullptr && nq <= 16 && !no_cluster && (kc == 64 || kc == 128 || kc == 256)) {
        if (kc == 64 && fp.dim <= 1024) launch_fin(scan_finalize_cluster_kernel<2, false, 8>, nq * kFinCluster, 256, st, pdl, fp);
        else if (kc == 64) launch_fin(scan_finalize_cluster_kernel<2>, nq * kFinCluster, 256, st, pdl, fp);
        else if (kc == 128 && fp.dim <= 1024) launch_fin(scan_finalize_cluster_kernel<4, false, 8>, nq * kFinCluster, 256, st, pdl, fp);
        else if (kc == 128) launch_fin(scan_finalize_cluster_kernel<4>, nq * kFinCluster, 256, st, pdl, fp);
        else launch_fin(scan_finalize_cluster_kernel<8>, nq * kFinCluster, 256, st, pdl, fp);
        CDR_LAUNCH_CHECK();
        return CDR_OK;
    }
    // Large batches of unsorted lists (the tensor-core lane: 1024 queries per batch): a 1024-thread CTA owns a whole SM
    // (64 registers x 1024 threads) and spends most of its 25 us waiting on its own merge tree and row reads, so 1024 of
    // them are 7 waves; 256-thread CTAs run 4 per SM and overlap each other's latencies.  CADENCE_FIN_WARPS=32 keeps the
    // wide form (A/B aid).  Same arithmetic and order => same bits.
    static const int fin_warps = [] { const char *e = getenv("CADENCE_FIN_WARPS"); return e ? atoi(e) : 8; }();
    if (kc == 128 && fp.counts != nullptr && nq >= 256 && fin_warps == 8) {
        scan_finalize_kernel<4, 8, 2><<<nq, 256, 0, st>>>(fp);
        CDR_LAUNCH_CHECK();
        return CDR_OK;
    }
    if (kc == 64) launch_fin(scan_finalize_kernel<2, 32, 2>, nq, 1024, st, pdl, fp);
    else if (kc == 128) launch_fin(scan_finalize_kernel<4, 32, 2>, nq, 1024, st, pdl, fp);
    else if (kc == 256) launch_fin(scan_finalize_kernel<8, 16, 4>, nq, 512, st, pdl, fp);
    else {
        cdr_set_error("finalize: candidate width %d not built", kc);
        return CDR_ERR_UNSUPPORTED;
    }
    CDR_LAUNCH_CHECK();
    return CDR_OK;
}

// q_index / q_count (QPC == 1 only): the conditional re-run of ScanParams -- nq is then the number of slots.
template <int J, int RPW, int NPL, int QPC, bool BF = false>
int launch_scan_t(cdr_store *s, ScanWorkspace &ws, const float *q_dev, int nq,
                  const uint32_t *allow, int k, double *out_score, int64_t *out_id, int32_t *out_n,
                  cudaStream_t st, const int *q_index = nullptr, const int *q_count = nullptr,
                  const ScanFinalizeOn *fin = nullptr, const PeerLink *peer = nullptr)
{
    using L = ScanSmem<J, RPW, NPL, QPC, BF>;
    constexpr int KC = L::KC;
    // per-device launch configuration (the smem opt-in attribute is per device)
    // (stores on the same device have different mutexes: the first-use initialisation takes its own lock)
    static std::mutex init_mu;
    static int stages_by_dev[64] = {0};
    static size_t smem_by_dev[64] = {0};
    int stages;
    size_t smem;
    {
        std::lock_guard<std::mutex> init_lock(init_mu);
        if (stages_by_dev[s->device & 63] == 0) {
            int dev_smem = 0;
            CDR_CUDA(cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, s->device));
            const int st_n = L::max_stages((size_t)dev_smem);
            CDR_CUDA(cudaFuncSetAttribute(exact_scan_kernel<J, RPW, NPL, QPC, BF>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::bytes(st_n)));
            smem_by_dev[s->device & 63] = L::bytes(st_n);
            stages_by_dev[s->device & 63] = st_n;
        }
        stages = stages_by_dev[s->device & 63];
        smem = smem_by_dev[s->device & 63];
    }
    const int64_t n_tiles = (s->n_rows + L::TR - 1) / L::TR;
    const int n_groups = (nq + QPC - 1) / QPC;
    // Shared reads with several query groups: the groups split the SMs instead of queueing behind each other with
    // a full-width grid each.  The bytes streamed are the same, but a CTA then sees n_groups times more rows per
    // query, so its running top-k saturates and far fewer rows pass the threshold test (1 M rows, 64 queries:
    // 27 % -> 5 % of the rows cost an insertion), and the finalize kernel merges n_groups times fewer lists.
    // CADENCE_K1_SPLIT=0 keeps a full-width grid per group (A/B aid).
    static const bool k1_split = [] { const char *e = getenv("CADENCE_K1_SPLIT"); return !(e && e[0] == '0'); }();
    int sm_share = s->sm_count;
    if (QPC > 1 && n_groups > 1 && k1_split) sm_share = s->sm_count / n_groups > 0 ? s->sm_count / n_groups : 1;
    int grid = (int)(n_tiles < sm_share ? n_tiles : sm_share);
    if (grid < 1) grid = 1;

    // Selective filters.  The reference's planner sends small scoped candidate sets (<= 2 000 rows by default,
    // app/retrieve.py:277-287) to the exact lane; scanning the whole table for them wastes the bus.  For a
    // caller-supplied bitmap the allowed rows are compacted into a list (one pass over n_rows/8 bytes) and a
    // GATHER launch scans only those rows when there are at most n_rows/2 of them (4 KB row gathers stream at
    // ~5.4 TB/s: 1/64 of the rows 0.046 ms, 1/4 0.221 ms, 1/2 0.405 ms vs 0.61 ms for the full scan of 1 M rows,
    // profiles/r01/k1_gather_sweep.json); otherwise it exits at once
    // and the full scan below serves.  The decision is taken on the device (no host round trip); the skipped
    // launch writes empty candidate lists.  Same per-row arithmetic => identical results either way.
    static const bool k1_nogather = [] { const char *e = getenv("CADENCE_K1_GATHER"); return e && e[0] == '0'; }();
    int g_gather = 0;
    uint32_t list_cap = 0;
    if (allow != nullptr && allow != s->valid && !k1_nogather && s->n_rows >= 4 * L::TR) {
        static const int64_t gather_div = [] {       // CADENCE_K1_GATHER_DIV: A/B aid, list capacity = rows / div
            const char *e = getenv("CADENCE_K1_GATHER_DIV");
            const int v = e ? atoi(e) : 2;
            return (int64_t)(v >= 1 ? v : 2);
        }();
        int64_t cap = s->n_rows / gather_div;
        if (cap < L::TR) cap = L::TR;
        list_cap = (uint32_t)cap;
        const int64_t gt = (cap + L::TR - 1) / L::TR;
        g_gather = (int)(gt < sm_share ? gt : sm_share);
        if (cdr_ws_reserve((void **)&ws.row_list, &ws.row_list_bytes, (size_t)cap * 4 + 16) != CDR_OK) return CDR_ERR_OOM;
        unsigned int *cnt = reinterpret_cast<unsigned int *>(ws.row_list) + cap;      // the count lives behind the list
        CDR_CUDA(cudaMemsetAsync(cnt, 0, 4, st));
        const int64_t words = (s->n_rows + 31) / 32;
        int64_t blocks = (words + 255) / 256;
        if (blocks > (int64_t)s->sm_count * 4) blocks = (int64_t)s->sm_count * 4;
        compact_allow_kernel<<<(unsigned)blocks, 256, 0, st>>>(allow, s->n_rows, ws.row_list, list_cap, cnt);
        CDR_LAUNCH_CHECK();
    }
    const int g_total = g_gather + grid;

    const size_t need = (size_t)nq * g_total * KC * sizeof(uint64_t);
    if (cdr_ws_reserve((void **)&ws.cta_keys, &ws.cta_keys_bytes, need) != CDR_OK) return CDR_ERR_OOM;
    // work-stealing counters: zeroed when (re)allocated, re-zeroed by every finalize launch
    static const bool k1_static = [] { const char *e = getenv("CADENCE_K1_SCHED"); return e && e[0] == 's'; }();
    if (!k1_static && ws.tile_ctr_bytes < (size_t)nq * 4) {
        if (cdr_ws_reserve((void **)&ws.tile_ctr, &ws.tile_ctr_bytes, (size_t)nq * 4) != CDR_OK) return CDR_ERR_OOM;
        CDR_CUDA(cudaMemsetAsync(ws.tile_ctr, 0, ws.tile_ctr_bytes, st));
    }

    ScanParams sp;
    sp.rows = BF ? reinterpret_cast<const float *>(s->emb_bf16) : s->emb_f32;
    sp.inv_norm = s->inv_norm;
    sp.allow = allow;
    sp.queries = q_dev;
    sp.nq = nq;
    sp.cta_keys = ws.cta_keys;
    sp.n_rows = s->n_rows;
    sp.n_tiles = n_tiles;
    sp.n_stages = stages;
    sp.tile_ctr = k1_static ? nullptr : ws.tile_ctr;
    sp.row_list = nullptr;
    sp.list_count = nullptr;
    sp.list_cap = list_cap;
    sp.lists_per_query = g_total;
    sp.list_offset = 0;
    sp.q_index = QPC == 1 ? q_index : nullptr;
    sp.q_count = QPC == 1 ? q_count : nullptr;
    // One scan per query (QPC == 1): the CTAs are persistent over the queries of the launch (gridDim.y = 1), so a CTA's
    // producer warp prefetches query y+1's first tiles while its consumers sort and merge the lists of query y, instead
    // of a fresh CTA per (SM, query) paying pipeline fill and epilogue alone.  CADENCE_K1_PERSIST=0 restores one CTA
    // per (SM, query) (A/B aid).  Shared-read launches keep one grid row per query group (the groups split the SMs).
    static const bool k1_persist = [] { const char *e = getenv("CADENCE_K1_PERSIST"); return !(e && e[0] == '0'); }();
    const int grid_y = (QPC == 1 && (k1_persist || q_index != nullptr)) ? 1 : n_groups;

    cdr_prof_mark_begin(0, st);
    if (g_gather > 0) {
        ScanParams gp = sp;
        gp.row_list = ws.row_list;
        gp.list_count = reinterpret_cast<unsigned int *>(ws.row_list) + list_cap;
        gp.allow = nullptr;
        exact_scan_kernel<J, RPW, NPL, QPC, BF><<<dim3(g_gather, grid_y), (L::CW + 1) * 32, smem, st>>>(gp);
        CDR_LAUNCH_CHECK();
        sp.list_count = gp.list_count;
        sp.list_offset = g_gather;
    }
    exact_scan_kernel<J, RPW, NPL, QPC, BF><<<dim3(grid, grid_y), (L::CW + 1) * 32, smem, st>>>(sp);
    CDR_LAUNCH_CHECK();
    cdr_prof_mark_end(0, st);

    FinalizeParams fp;
    fp.lists = ws.cta_keys;
    fp.n_lists = g_total;
    fp.counts = nullptr;
    fp.cap = 0;
    fp.rows = s->emb_f32;                            // BF: exact re-score on the fp32 rows when they are resident
    fp.bf16_rows = BF ? s->emb_bf16 : nullptr;
    fp.queries = q_dev;
    fp.ids = s->ids;
    fp.dim = s->dim;
    fp.k = k;
    fp.out_score = out_score;
    fp.out_id = out_id;
    fp.out_n = out_n;
    fp.reset_ctr = sp.tile_ctr;
    fp.reset_div = QPC;
    fp.q_index = sp.q_index;
    fp.q_count = sp.q_count;
    fp.n_slots = nq;
    if (peer != nullptr) fp.peer = *peer;
    cudaStream_t fst = st;
    if (fin != nullptr && fin->stream != st) {          // the finalize (and what follows it) runs beside the next scan
        CDR_CUDA(cudaEventRecord(fin->scan_done, st));
        CDR_CUDA(cudaStreamWaitEvent(fin->stream, fin->scan_done, 0));
        fst = fin->stream;
    }
    // Programmatic dependent launch of the finalize behind the scan on the same stream (CADENCE_PDL=0: plain stream order;
    // also off while the profiling marks record an event between the two kernels).
    static const bool pdl_on = [] { const char *e = getenv("CADENCE_PDL"); return !(e && e[0] == '0'); }();
    const bool pdl = pdl_on && fst == st && !cdr_prof_active();
#ifdef CDR_FIN_TIMING
    // probe: the same finalize twice -- the second run finds its code and rows in the caches
    static const bool fin_twice = [] { const char *e = getenv("CADENCE_FIN_TWICE"); return e && e[0] == '1'; }();
    if (fin_twice && peer == nullptr) launch_finalize(fp, KC, sp.q_index != nullptr && nq > 64 ? 64 : nq, fst, pdl);
#endif
    return launch_finalize(fp, KC, sp.q_index != nullptr && nq > 64 ? 64 : nq, fst, pdl);
}

enum ScanMode { kScanSingle = 0, kScanShared = 1, kScanDeep = 2 };

template <int NPL>
int launch_scan_dim(cdr_store *s, ScanWorkspace &ws, const float *q_dev, int nq,
                    const uint32_t *allow, int k, double *out_score, int64_t *out_id,
                    int32_t *out_n, cudaStream_t st, ScanMode mode, const int *q_index = nullptr,
                    const int *q_count = nullptr, const ScanFinalizeOn *fin = nullptr, const PeerLink *peer = nullptr)
{
#define CDR_SCAN_CASE(J_, RPW_)                                                                              \
    if constexpr (NPL == 2) {                                                                                \
        if (mode == kScanDeep)                                                                               \
            return launch_scan_t<J_, kDeepRPW, NPL, kDeepQPC>(s, ws, q_dev, nq, allow, k, out_score, out_id, out_n, st, \
                                                              nullptr, nullptr, nullptr, peer);              \
    }                                                                                                        \
    if (mode != kScanSingle)                                                                                 \
        return launch_scan_t<J_, RPW_, NPL, kSharedQPC>(s, ws, q_dev, nq, allow, k, out_score, out_id, out_n, st, nullptr, \
                                                        nullptr, nullptr, peer);                             \
    return launch_scan_t<J_, RPW_, NPL, 1>(s, ws, q_dev, nq, allow, k, out_score, out_id, out_n, st, q_index, q_count, fin, peer)
    switch (s->dim) {
    case 256:  CDR_SCAN_CASE(2, 2);
    case 512:  CDR_SCAN_CASE(4, 2);
    case 768:  CDR_SCAN_CASE(6, 2);
    case 1024: CDR_SCAN_CASE(8, 2);
    // kSharedQPC queries x J float4 would not fit the register budget: one scan per query
    case 1536: return launch_scan_t<12, 1, NPL, 1>(s, ws, q_dev, nq, allow, k, out_score, out_id, out_n, st, q_index, q_count, fin, peer);
    case 2048: return launch_scan_t<16, 1, NPL, 1>(s, ws, q_dev, nq, allow, k, out_score, out_id, out_n, st, q_index, q_count, fin, peer);
    default:
        cdr_set_error("exact scan: dim %d not built (supported: 256,512,768,1024,1536,2048)", s->dim);
        return CDR_ERR_UNSUPPORTED;
    }
#undef CDR_SCAN_CASE
}

}  // namespace

// Chooses the candidate-list width: KC = 64 serves k <= 56, KC = 256 serves k <= 248
// (KC - k >= 8 spare slots absorb fp32-vs-fp64 rank swaps at the boundary) -- and, for shared reads, the kernel:
// whole groups of kDeepQPC = 16 queries take the deep kernel (k <= 56, dim <= 1024), a tail of >= 10 queries is
// padded into one more deep group, a shorter tail (and batches of <= 9) runs as groups of kSharedQPC = 3 in a second
// launch.  Measured at 1 M rows (profiles/r01/k1_shared_probe_*.json): a deep group costs 1.6-2.1 ms whatever its
// fill, three queries in registers 0.63 ms, so 17 queries are 2.05 + 0.57 ms instead of 4.3 ms as two deep groups.
// CADENCE_K1_DEEP=0 keeps the 3-queries-in-registers kernel for every batch size (A/B aid).
int cdr_exact_scan_launch(cdr_store *s, ScanWorkspace &ws, const float *q_dev, int nq,
                          const uint32_t *allow, int k, double *out_score, int64_t *out_id,
                          int32_t *out_n, cudaStream_t st, bool share_reads, const ScanFinalizeOn *fin, const PeerLink *peer)
{
    const bool share = share_reads && nq >= 2;
    if (fin != nullptr && !share)       // (one scan per query only: the pipelined sharded step)
        return k > 56 ? launch_scan_dim<8>(s, ws, q_dev, nq, allow, k, out_score, out_id, out_n, st, kScanSingle, nullptr, nullptr, fin)
                      : launch_scan_dim<2>(s, ws, q_dev, nq, allow, k, out_score, out_id, out_n, st, kScanSingle, nullptr, nullptr, fin);
    if (k > 56)
        return launch_scan_dim<8>(s, ws, q_dev, nq, allow, k, out_score, out_id, out_n, st, share ? kScanShared : kScanSingle,
                                  nullptr, nullptr, nullptr, peer);
    static const bool k1_deep = [] { const char *e = getenv("CADENCE_K1_DEEP"); return !(e && e[0] == '0'); }();
    int n_deep = 0;
    if (share && k1_deep && s->dim <= 1024) {
        n_deep = nq / kDeepQPC * kDeepQPC;
        if (nq - n_deep >= 10) n_deep = nq;
    }
    if (n_deep > 0) {
        const int rc = launch_scan_dim<2>(s, ws, q_dev, n_deep, allow, k, out_score, out_id, out_n, st, kScanDeep, nullptr,
                                          nullptr, nullptr, peer);
        if (rc != CDR_OK || n_deep == nq) return rc;
    }
    const int rest = nq - n_deep;
    PeerLink tail;                       // the second launch's queries sit behind the deep groups' in the exchange buffers
    if (peer != nullptr) {
        tail = *peer;
        tail.q0 += n_deep;
    }
    return launch_scan_dim<2>(s, ws, q_dev + (size_t)n_deep * s->dim, rest, allow, k, out_score + (size_t)n_deep * k,
                              out_id + (size_t)n_deep * k, out_n + n_deep, st, share && rest >= 2 ? kScanShared : kScanSingle,
                              nullptr, nullptr, nullptr, peer != nullptr ? &tail : nullptr);
}

// The scan lanes' finalize kernels can end with the exchange (PeerLink) for these shapes: candidate width 64 or 128
// (exact lane: k <= 56; bf16-row lane: k <= 120), merge scratch in static shared memory.
bool cdr_scan_peer_fusable(bool bf16_rows, int k, int world)
{
    return world >= 2 && world <= kPeerFusedRanks && k <= kPeerFusedK && k <= (bf16_rows ? 120 : 56);
}

// Conditional re-run of the queries q_index[0 .. min(*q_count, n_slots)) of a batch on the exact lane, one scan per
// query, results written to the queries' own output rows (K2's overflow path: the count lives on the device, and a
// zero count costs two empty launches).  Store mutex held by the caller.
int cdr_exact_scan_redo_launch(cdr_store *s, ScanWorkspace &ws, const float *q_dev, int n_slots, const uint32_t *allow,
                               int k, double *out_score, int64_t *out_id, int32_t *out_n, cudaStream_t st,
                               const int *q_index, const int *q_count)
{
    if (k > 56) return launch_scan_dim<8>(s, ws, q_dev, n_slots, allow, k, out_score, out_id, out_n, st, kScanSingle, q_index, q_count);
    return launch_scan_dim<2>(s, ws, q_dev, n_slots, allow, k, out_score, out_id, out_n, st, kScanSingle, q_index, q_count);
}

// The "ann" scan lane: the same scan over the bf16 copy of the rows (half the bytes), candidate lists twice as wide
// as the exact lane's (KC = 128 for k <= 120, else 256), exact fp64 re-score of the survivors.  One query: one pass.  A
// batch: its queries SHARE passes in pairs (every 16-byte shared-memory load feeds two queries, the 2 x 4 dot products
// of a warp are reduced together): 0.39 ms per pair over 1 M rows against 0.30 ms for one query; four queries per pass
// with 2 rows per warp turned issue-bound (0.79 ms) and is not built (profiles/r02/bf16_share_probe.jsonl).
// CADENCE_BF16_SHARE=0: no sharing (A/B aid).  Same per-(row, query) arithmetic and reduction => the same bits.
int cdr_bf16_scan_launch(cdr_store *s, ScanWorkspace &ws, const float *q_dev, int nq, const uint32_t *allow, int k,
                         double *out_score, int64_t *out_id, int32_t *out_n, cudaStream_t st, const int *q_index,
                         const int *q_count, const PeerLink *peer)
{
    static const int share_env = [] { const char *e = getenv("CADENCE_BF16_SHARE"); return e ? atoi(e) : -1; }();
    int share = 1;
    if (q_index == nullptr && nq >= 2 && k <= 120 && CDR_BF16_FHFMA) share = share_env == 0 ? 1 : 2;
#define CDR_BF_CASE(J_)                                                                                          \
    if (share == 2) return launch_scan_t<J_, 4, 4, 2, true>(s, ws, q_dev, nq, allow, k, out_score, out_id, out_n, st, nullptr, nullptr, nullptr, peer); \
    if (k <= 120) return launch_scan_t<J_, 4, 4, 1, true>(s, ws, q_dev, nq, allow, k, out_score, out_id, out_n, st, q_index, q_count, nullptr, peer); \
    return launch_scan_t<J_, 4, 8, 1, true>(s, ws, q_dev, nq, allow, k, out_score, out_id, out_n, st, q_index, q_count, nullptr, peer)
    switch (s->dim) {
    case 256:  CDR_BF_CASE(2);
    case 512:  CDR_BF_CASE(4);
    case 768:  CDR_BF_CASE(6);
    case 1024: CDR_BF_CASE(8);
    default:
        cdr_set_error("bf16 scan: dim %d not built (supported: 256,512,768,1024)", s->dim);
        return CDR_ERR_UNSUPPORTED;
    }
#undef CDR_BF_CASE
}

// Used by the batched bf16 lane (gemm_topk.cu): select the top-kc of one unsorted candidate
// list per query, re-score them exactly and order them.
int cdr_finalize_unsorted_launch(cdr_store *s, const uint64_t *lists, const uint32_t *counts, int cap,
                                 int kc, const float *q_dev, int nq, int k, bool use_bf16_rows,
                                 double *out_score, int64_t *out_id, int32_t *out_n, cudaStream_t st)
{
    FinalizeParams fp;
    fp.lists = lists;
    fp.n_lists = 1;
    fp.counts = counts;
    fp.cap = cap;
    fp.rows = use_bf16_rows ? nullptr : s->emb_f32;
    fp.bf16_rows = s->emb_bf16;
    fp.queries = q_dev;
    fp.ids = s->ids;
    fp.dim = s->dim;
    fp.k = k;
    fp.out_score = out_score;
    fp.out_id = out_id;
    fp.out_n = out_n;
    fp.reset_ctr = nullptr;
    fp.reset_div = 1;
    fp.q_index = nullptr;
    fp.q_count = nullptr;
    fp.n_slots = nq;
    return launch_finalize(fp, kc, nq, st);
}
