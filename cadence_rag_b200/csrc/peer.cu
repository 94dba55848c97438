// peer.cu -- K4p: the multi-GPU exchange step fused with the k-way merge, over NVLink peer memory.
//
// Row-sharded corpus, one process per GPU.  The baseline exchange is an NCCL
// all-gather of every rank's [nq,k] (score,id,n) lists followed by the K4 merge kernel
// (topk_merge.cu) -- two launches plus the pack/unpack copies around the collective.  Here each
// rank owns one receive buffer that every peer maps with CUDA IPC, and ONE kernel per rank does
//   push   : CTA q stores this rank's list for query q into slot[parity][me][q] of EVERY rank's
//            buffer (posted NVLink writes), then publishes flag[me][q] = epoch with a
//            system-scope release store;
//   wait   : R threads spin (acquire loads on LOCAL memory) until flag[r][q] >= epoch for every r;
//   merge  : the R lists of query q are ranked by counting with the global ordering rule
//            (score desc, NaN last, id asc) and the first k written out.
// All ranks end with the identical result, as with the all-gather.  No CTA waits before it has
// pushed, so the protocol cannot deadlock however the grids are scheduled.
//
// Slots are double-buffered by epoch parity: rank A can only finish epoch e+1 after every peer has
// STARTED epoch e+1, i.e. after that peer's epoch-e kernel (which read parity e&1) completed, so A's
// epoch e+2 writes never land on data a peer still reads.  All calls of a group must be issued in the
// same order on every rank and on one stream per rank.
#include "common.cuh"

#include <mutex>

#include <cstdlib>
#include <cstring>

struct cdr_peer_group {
    int device = 0, rank = 0, world = 1;
    int max_nq = 0, max_k = 0;
    size_t entry_bytes = 0;       // one list: scores f64[max_k], ids i64[max_k], n i32 (+pad)
    size_t flags_off = 0;         // byte offset of flags u32[world][max_nq]
    size_t bytes = 0;
    unsigned char *local = nullptr;
    unsigned char *peer[CDR_PEER_MAX_RANKS] = {nullptr};   // mapped base of every rank's buffer (own included)
    bool connected = false;
    uint32_t epoch = 0;
    // local lists of cdr_search_sharded (this rank's [max_nq, max_k] results between the lane and the exchange)
    double *loc_score = nullptr;
    int64_t *loc_id = nullptr;
    int32_t *loc_n = nullptr;
    // pipelined step (cdr_search_sharded, one scan per query, batches of more than kPipeChunk queries): the finalize and
    // exchange kernels of chunk c run on `side` while the scan of chunk c+1 runs on the caller's stream
    cudaStream_t side = nullptr;
    cudaEvent_t scan_done[2] = {nullptr, nullptr};
    cudaEvent_t fin_done[2] = {nullptr, nullptr};
};

namespace {

struct PeerParams {
    PeerLink link;
    const double *scores;   // this rank's lists [nq, k]
    const int64_t *ids;
    const int32_t *n;
    int nq, k;
    double *out_score;
    int64_t *out_id;
    int32_t *out_n;
};

__global__ void __launch_bounds__(256) peer_publish_merge_kernel(const PeerParams p)
{
    extern __shared__ unsigned char raw[];
    uint64_t *s_sc = reinterpret_cast<uint64_t *>(raw);                    // order keys of the scores
    int64_t *s_id = reinterpret_cast<int64_t *>(s_sc + (size_t)p.link.world * p.k);
    __shared__ int s_cnt;
    __shared__ int s_n[CDR_PEER_MAX_RANKS];
    const int q = blockIdx.x;
    peer_exchange_cta(p.link, p.link.q0 + q, p.k, p.scores + (size_t)q * p.k, p.ids + (size_t)q * p.k, p.n[q], s_sc, s_id,
                      s_n, &s_cnt, p.out_score + (size_t)q * p.k, p.out_id + (size_t)q * p.k, p.out_n + q);
}

}  // namespace

extern "C" int32_t cdr_peer_group_create(cdr_peer_group **out, int32_t device, int32_t rank, int32_t world,
                                         int32_t max_nq, int32_t max_k, void *out_handle)
{
    CDR_REQUIRE(out && out_handle, CDR_ERR_INVALID, "cdr_peer_group_create: NULL argument");
    CDR_REQUIRE(world >= 1 && world <= CDR_PEER_MAX_RANKS && rank >= 0 && rank < world, CDR_ERR_INVALID,
                "cdr_peer_group_create: need 1 <= world <= %d and 0 <= rank < world", CDR_PEER_MAX_RANKS);
    CDR_REQUIRE(max_nq >= 1 && max_k >= 1 && max_k <= CDR_MAX_K && (int64_t)world * max_k <= 4096, CDR_ERR_INVALID,
                "cdr_peer_group_create: need max_nq >= 1, 1 <= max_k <= %d, world*max_k <= 4096", CDR_MAX_K);
    static_assert(sizeof(cudaIpcMemHandle_t) == CDR_PEER_HANDLE_BYTES, "IPC handle size");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
        cdr_set_error("cdr_peer_group_create: CUDA device %d not available; this engine has no CPU fallback", device);
        return CDR_ERR_NO_DEVICE;
    }
    DeviceGuard g(device);
    cdr_peer_group *pg = new cdr_peer_group();
    pg->device = device; pg->rank = rank; pg->world = world; pg->max_nq = max_nq; pg->max_k = max_k;
    pg->entry_bytes = (size_t)max_k * 16 + 16;
    pg->flags_off = ((size_t)2 * world * max_nq * pg->entry_bytes + 255) & ~(size_t)255;
    pg->bytes = pg->flags_off + (size_t)world * max_nq * 4;
    cudaError_t e = cudaMalloc(&pg->local, pg->bytes);
    if (e == cudaSuccess) e = cudaMemset(pg->local, 0, pg->bytes);
    if (e == cudaSuccess) e = cudaMalloc(&pg->loc_score, (size_t)max_nq * max_k * 8);
    if (e == cudaSuccess) e = cudaMalloc(&pg->loc_id, (size_t)max_nq * max_k * 8);
    if (e == cudaSuccess) e = cudaMalloc(&pg->loc_n, (size_t)max_nq * 4);
    if (e == cudaSuccess) {
        int lo = 0, hi = 0;                                   // the side stream's small kernels go first when SMs free up
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        e = cudaStreamCreateWithPriority(&pg->side, cudaStreamNonBlocking, hi);
    }
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
        e = cudaEventCreateWithFlags(&pg->scan_done[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&pg->fin_done[i], cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, pg->local);
    if (e != cudaSuccess) {
        cdr_set_error("cdr_peer_group_create: %s", cudaGetErrorString(e));
        cudaFree(pg->local); cudaFree(pg->loc_score); cudaFree(pg->loc_id); cudaFree(pg->loc_n);
        delete pg;
        return e == cudaErrorMemoryAllocation ? CDR_ERR_OOM : CDR_ERR_CUDA;
    }
    memcpy(out_handle, &h, sizeof(h));
    pg->peer[rank] = pg->local;
    *out = pg;
    return CDR_OK;
}

extern "C" int32_t cdr_peer_group_connect(cdr_peer_group *pg, const void *all_handles)
{
    CDR_REQUIRE(pg && all_handles, CDR_ERR_INVALID, "cdr_peer_group_connect: NULL argument");
    CDR_REQUIRE(!pg->connected, CDR_ERR_STATE, "cdr_peer_group_connect: already connected");
    DeviceGuard g(pg->device);
    for (int r = 0; r < pg->world; ++r) {
        if (r == pg->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const unsigned char *)all_handles + (size_t)r * CDR_PEER_HANDLE_BYTES, sizeof(h));
        void *ptr = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cdr_set_error("cdr_peer_group_connect: cannot map rank %d's buffer (%s); peer memory needs all ranks on "
                          "one node with P2P access", r, cudaGetErrorString(e));
            cudaGetLastError();
            for (int o = 0; o < r; ++o)
                if (o != pg->rank && pg->peer[o]) { cudaIpcCloseMemHandle(pg->peer[o]); pg->peer[o] = nullptr; }
            return CDR_ERR_UNSUPPORTED;
        }
        pg->peer[r] = (unsigned char *)ptr;
    }
    pg->connected = true;
    return CDR_OK;
}

extern "C" int32_t cdr_peer_group_destroy(cdr_peer_group *pg)
{
    if (!pg) return CDR_OK;
    DeviceGuard g(pg->device);
    cudaDeviceSynchronize();
    for (int r = 0; r < pg->world; ++r)
        if (r != pg->rank && pg->peer[r]) cudaIpcCloseMemHandle(pg->peer[r]);
    cudaFree(pg->local); cudaFree(pg->loc_score); cudaFree(pg->loc_id); cudaFree(pg->loc_n);
    for (int i = 0; i < 2; ++i) {
        if (pg->scan_done[i]) cudaEventDestroy(pg->scan_done[i]);
        if (pg->fin_done[i]) cudaEventDestroy(pg->fin_done[i]);
    }
    if (pg->side) cudaStreamDestroy(pg->side);
    delete pg;
    return CDR_OK;
}

// The group's view for its NEXT epoch (one epoch = one exchange of up to max_nq queries).
static PeerLink next_epoch_link(cdr_peer_group *pg)
{
    PeerLink l;
    for (int r = 0; r < CDR_PEER_MAX_RANKS; ++r) l.base[r] = pg->peer[r];
    l.rank = pg->rank; l.world = pg->world; l.max_nq = pg->max_nq; l.max_k = pg->max_k;
    l.entry_bytes = pg->entry_bytes; l.flags_off = pg->flags_off;
    l.epoch = ++pg->epoch;
    l.q0 = 0;
    return l;
}

// The pipelined form of a sharded step (one scan per query, more than kPipeChunk queries): the batch goes through in chunks
// of kPipeChunk queries -- the size the 8-CTA-cluster finalize serves; its 256-thread CTAs (80 registers, 10 KB) and the
// exchange kernel's fit on an SM BESIDE a scan CTA (288 threads, 126 registers, 197 KB) -- with the finalize + exchange of chunk
// c on the group's side stream while the scan of chunk c+1 runs on the caller's stream.  Only the last chunk's tail (and the
// wait for the slowest rank inside its exchange) is left on the critical path; the call still completes in stream order on
// the caller's stream.  Two scan workspaces alternate; a workspace is re-used only after its chunk's finalize has finished.
constexpr int kPipeChunk = 16;

static int sharded_exact_pipelined(cdr_store *s, cdr_peer_group *pg, const float *q_dev, int nq, int k,
                                   const uint32_t *allow_dev, double *out_score_dev, int64_t *out_id_dev,
                                   int32_t *out_n_dev, cudaStream_t st)
{
    DeviceGuard g(s->device);
    std::lock_guard<std::mutex> lk(s->mu);
    const uint32_t *allow = allow_dev ? allow_dev : (s->any_invalid ? s->valid : nullptr);
    int last = -1;
    for (int q0 = 0, c = 0; q0 < nq; q0 += kPipeChunk, ++c) {
        const int m = nq - q0 < kPipeChunk ? nq - q0 : kPipeChunk;
        const int par = c & 1;
        const int l0 = q0 % pg->max_nq;                          // slice of the group's local-list buffers
        if (c >= 2) CDR_CUDA(cudaStreamWaitEvent(st, pg->fin_done[par], 0));     // this workspace's previous chunk is done
        ScanFinalizeOn fin{pg->side, pg->scan_done[par]};
        int rc = cdr_exact_scan_launch(s, s->ws_pipe[par][st], q_dev + (size_t)q0 * s->dim, m, allow, k,
                                       pg->loc_score + (size_t)l0 * k, pg->loc_id + (size_t)l0 * k, pg->loc_n + l0, st,
                                       /*share_reads=*/false, &fin);
        if (rc != CDR_OK) return rc;
        rc = cdr_peer_exchange_merge(pg, pg->loc_score + (size_t)l0 * k, pg->loc_id + (size_t)l0 * k, pg->loc_n + l0, m, k,
                                     out_score_dev + (size_t)q0 * k, out_id_dev + (size_t)q0 * k, out_n_dev + q0, pg->side);
        if (rc != CDR_OK) return rc;
        CDR_CUDA(cudaEventRecord(pg->fin_done[par], pg->side));
        last = par;
    }
    if (last >= 0) CDR_CUDA(cudaStreamWaitEvent(st, pg->fin_done[last], 0));     // the side stream is in order: the last event covers all
    return CDR_OK;
}

extern "C" int32_t cdr_peer_exchange_merge(cdr_peer_group *pg, const double *scores_dev, const int64_t *ids_dev,
                                           const int32_t *n_dev, int32_t nq, int32_t k, double *out_score_dev,
                                           int64_t *out_id_dev, int32_t *out_n_dev, void *stream)
{
    CDR_REQUIRE(pg && scores_dev && ids_dev && n_dev && out_score_dev && out_id_dev && out_n_dev, CDR_ERR_INVALID,
                "cdr_peer_exchange_merge: NULL argument");
    CDR_REQUIRE(pg->connected || pg->world == 1, CDR_ERR_STATE, "cdr_peer_exchange_merge: group not connected");
    CDR_REQUIRE(nq >= 0 && k >= 1 && k <= pg->max_k, CDR_ERR_INVALID, "cdr_peer_exchange_merge: k=%d outside [1,%d]", k,
                pg->max_k);
    DeviceGuard g(pg->device);
    const size_t smem = (size_t)pg->world * k * 16;
    {
        static std::mutex attr_mu;
        static bool attr_set[64] = {false};
        std::lock_guard<std::mutex> attr_lock(attr_mu);
        if (!attr_set[pg->device & 63] && smem > 48 * 1024) {
            CDR_CUDA(cudaFuncSetAttribute(peer_publish_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096 * 16));
            attr_set[pg->device & 63] = true;
        }
    }
    for (int q0 = 0; q0 < nq; q0 += pg->max_nq) {
        const int m = nq - q0 < pg->max_nq ? nq - q0 : pg->max_nq;
        PeerParams p;
        p.link = next_epoch_link(pg);
        p.scores = scores_dev + (size_t)q0 * k; p.ids = ids_dev + (size_t)q0 * k; p.n = n_dev + q0;
        p.nq = m; p.k = k;
        p.out_score = out_score_dev + (size_t)q0 * k; p.out_id = out_id_dev + (size_t)q0 * k; p.out_n = out_n_dev + q0;
        peer_publish_merge_kernel<<<m, 256, smem, (cudaStream_t)stream>>>(p);
        CDR_LAUNCH_CHECK();
    }
    return CDR_OK;
}

extern "C" int32_t cdr_search_sharded(cdr_store *s, cdr_peer_group *pg, int32_t lane, const float *q_dev, int32_t nq,
                                      int32_t k, const uint32_t *allow_dev, double *out_score_dev, int64_t *out_id_dev,
                                      int32_t *out_n_dev, void *stream)
{
    typedef int32_t (*lane_fn)(cdr_store *, const float *, int32_t, int32_t, const uint32_t *, double *, int64_t *,
                               int32_t *, void *);
    lane_fn fn = nullptr;
    switch (lane) {
    case CDR_DENSE_LANE_EXACT_F32: fn = cdr_search_exact_f32; break;
    case CDR_DENSE_LANE_EXACT_F32_SHARED: fn = cdr_search_exact_f32_shared; break;
    case CDR_DENSE_LANE_SCAN_BF16: fn = cdr_search_scan_bf16; break;
    case CDR_DENSE_LANE_BATCH_BF16: fn = cdr_search_batch_bf16; break;
    default: break;
    }
    CDR_REQUIRE(fn != nullptr, CDR_ERR_INVALID, "cdr_search_sharded: unknown lane %d", lane);
    CDR_REQUIRE(s != nullptr, CDR_ERR_INVALID, "cdr_search_sharded: store is NULL");
    if (pg == nullptr || pg->world == 1)
        return fn(s, q_dev, nq, k, allow_dev, out_score_dev, out_id_dev, out_n_dev, stream);
    CDR_REQUIRE(pg->connected, CDR_ERR_STATE, "cdr_search_sharded: group not connected");
    CDR_REQUIRE(pg->device == s->device, CDR_ERR_INVALID, "cdr_search_sharded: store on device %d, group on %d", s->device,
                pg->device);
    CDR_REQUIRE(k >= 1 && k <= pg->max_k, CDR_ERR_INVALID, "cdr_search_sharded: k=%d outside [1,%d]", k, pg->max_k);
    CDR_REQUIRE(nq >= 0 && out_score_dev && out_id_dev && out_n_dev, CDR_ERR_INVALID, "cdr_search_sharded: bad arguments");
    // Off by default: measured at 8 GPUs it LOSES (13 420 vs 14 160 queries/s, profiles/r02/README.md) -- four exchanges per
    // step are four points where every rank waits for the slowest one, and the scan of chunk c+2 waits for them through
    // its workspace; at 2 GPUs it is a wash.  CADENCE_SHARD_PIPELINE=1 turns it on (A/B aid, covered by the tests).
    static const bool pipelined = [] { const char *e = getenv("CADENCE_SHARD_PIPELINE"); return e && e[0] == '1'; }();
    if (pipelined && lane == CDR_DENSE_LANE_EXACT_F32 && nq > kPipeChunk && k <= 56 && pg->max_nq >= kPipeChunk &&
        pg->max_nq % kPipeChunk == 0 && s->finalized && s->emb_f32 != nullptr && q_dev != nullptr)
        return sharded_exact_pipelined(s, pg, q_dev, nq, k, allow_dev, out_score_dev, out_id_dev, out_n_dev,
                                       (cudaStream_t)stream);
    // Scan lanes: the exchange is folded into the lane's finalize kernel (the CTA that orders a query's list pushes it to the
    // peers, waits for theirs and merges: one launch and one local round trip less per step).  CADENCE_PEER_FUSED=0 keeps
    // the separate K4p launch (A/B aid; same bits, covered by the tests).
    static const bool fused_on = [] { const char *e = getenv("CADENCE_PEER_FUSED"); return !(e && e[0] == '0'); }();
    const bool bf_lane = lane == CDR_DENSE_LANE_SCAN_BF16;
    if (fused_on && lane != CDR_DENSE_LANE_BATCH_BF16 && cdr_scan_peer_fusable(bf_lane, k, pg->world)) {
        for (int q0 = 0; q0 < nq; q0 += pg->max_nq) {
            const int m = nq - q0 < pg->max_nq ? nq - q0 : pg->max_nq;
            const PeerLink link = next_epoch_link(pg);
            const int rc = bf_lane ? cdr_search_scan_bf16_peer(s, q_dev + (size_t)q0 * s->dim, m, k, allow_dev,
                                                               out_score_dev + (size_t)q0 * k, out_id_dev + (size_t)q0 * k,
                                                               out_n_dev + q0, stream, &link)
                                   : cdr_search_exact_peer(lane == CDR_DENSE_LANE_EXACT_F32_SHARED, s,
                                                           q_dev + (size_t)q0 * s->dim, m, k, allow_dev,
                                                           out_score_dev + (size_t)q0 * k, out_id_dev + (size_t)q0 * k,
                                                           out_n_dev + q0, stream, &link);
            // (an epoch consumed by a failed call is harmless only if every rank fails alike: argument errors are)
            if (rc != CDR_OK) return rc;
        }
        return CDR_OK;
    }
    for (int q0 = 0; q0 < nq; q0 += pg->max_nq) {
        const int m = nq - q0 < pg->max_nq ? nq - q0 : pg->max_nq;
        int rc = fn(s, q_dev + (size_t)q0 * s->dim, m, k, allow_dev, pg->loc_score, pg->loc_id, pg->loc_n, stream);
        if (rc != CDR_OK) return rc;
        rc = cdr_peer_exchange_merge(pg, pg->loc_score, pg->loc_id, pg->loc_n, m, k, out_score_dev + (size_t)q0 * k,
                                     out_id_dev + (size_t)q0 * k, out_n_dev + q0, stream);
        if (rc != CDR_OK) return rc;
    }
    return CDR_OK;
}
