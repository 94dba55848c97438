// abi.cu -- extern "C" surface glue: thread-local error text, launch counter, profiling
// events, and the exact-lane entry points (device- and host-buffer flavours).
#include "common.cuh"

#include <cstring>
#include <string>

std::atomic<int64_t> g_cdr_launches{0};

static thread_local char t_err[512] = "";

void cdr_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof(t_err), fmt, ap);
    va_end(ap);
}

extern "C" const char *cdr_last_error(void) { return t_err; }
extern "C" int32_t cdr_abi_version(void) { return CDR_ABI_VERSION; }
extern "C" int64_t cdr_kernel_launch_count(void) { return g_cdr_launches.load(); }

extern "C" int32_t cdr_device_count(int32_t *out_count)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (out_count) *out_count = (e == cudaSuccess) ? n : 0;
    if (e != cudaSuccess || n <= 0) {
        cdr_set_error("no CUDA device visible (%s); this engine has no CPU fallback",
                      e == cudaSuccess ? "count is 0" : cudaGetErrorString(e));
        return CDR_ERR_NO_DEVICE;
    }
    return CDR_OK;
}

// ---------------------------------------------------------------------------- per-thread pinned staging
namespace {
struct PinnedStage {
    void *p = nullptr;
    size_t n = 0;
};
thread_local PinnedStage t_stage;   // deliberately not freed at thread exit (the CUDA context may be gone by then)
}  // namespace

void *cdr_thread_pinned(size_t need)
{
    if (t_stage.n >= need && t_stage.p) return t_stage.p;
    if (t_stage.p) cudaFreeHost(t_stage.p);      // the thread's previous call has synchronised: not in use
    t_stage.p = nullptr;
    t_stage.n = 0;
    const size_t sz = need < 65536 ? 65536 : need * 2;
    cudaError_t e = cudaHostAlloc(&t_stage.p, sz, cudaHostAllocDefault);
    if (e != cudaSuccess) {
        cdr_set_error("cudaHostAlloc(%zu) for request staging failed: %s", sz, cudaGetErrorString(e));
        t_stage.p = nullptr;
        return nullptr;
    }
    t_stage.n = sz;
    return t_stage.p;
}

void *cdr_thread_device(int device, size_t need)
{
    thread_local std::map<int, PinnedStage> t_dev;      // same bookkeeping, device memory
    PinnedStage &b = t_dev[device];
    if (b.n >= need && b.p) return b.p;
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.n = 0;
    const size_t sz = need < 65536 ? 65536 : need + need / 2;
    cudaError_t e = cudaMalloc(&b.p, sz);
    if (e != cudaSuccess) {
        cdr_set_error("cudaMalloc(%zu) for request staging failed: %s", sz, cudaGetErrorString(e));
        b.p = nullptr;
        return nullptr;
    }
    b.n = sz;
    return b.p;
}

// ---------------------------------------------------------------------------- profiling
// When enabled, every launch of a dominant kernel (kind 0 = K1 exact scan, 1 = K2 batched
// bf16 GEMM) is bracketed by CUDA events recorded on the launching stream; cdr_prof_read
// synchronises those events and returns the summed device time.  Off by default (no events).
namespace {
struct ProfState {
    std::mutex mu;
    bool on = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pairs[2];
    cudaEvent_t pending[2] = {nullptr, nullptr};
};
ProfState g_prof;
}  // namespace

void cdr_prof_mark_begin(int kind, cudaStream_t st)
{
    if (!g_prof.on) return;
    std::lock_guard<std::mutex> lk(g_prof.mu);
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, st);
    g_prof.pending[kind] = e;
}

bool cdr_prof_active() { return g_prof.on; }

void cdr_prof_mark_end(int kind, cudaStream_t st)
{
    if (!g_prof.on) return;
    std::lock_guard<std::mutex> lk(g_prof.mu);
    if (!g_prof.pending[kind]) return;
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, st);
    g_prof.pairs[kind].push_back({g_prof.pending[kind], e});
    g_prof.pending[kind] = nullptr;
}

extern "C" int32_t cdr_prof_enable(int32_t on)
{
    std::lock_guard<std::mutex> lk(g_prof.mu);
    for (int k = 0; k < 2; ++k) {
        for (auto &pr : g_prof.pairs[k]) {
            cudaEventDestroy(pr.first);
            cudaEventDestroy(pr.second);
        }
        g_prof.pairs[k].clear();
        if (g_prof.pending[k]) cudaEventDestroy(g_prof.pending[k]);
        g_prof.pending[k] = nullptr;
    }
    g_prof.on = on != 0;
    return CDR_OK;
}

extern "C" int32_t cdr_prof_read(int32_t kind, double *out_total_ms, int64_t *out_launches)
{
    CDR_REQUIRE(kind == 0 || kind == 1, CDR_ERR_INVALID, "cdr_prof_read: kind must be 0 or 1");
    std::lock_guard<std::mutex> lk(g_prof.mu);
    double tot = 0.0;
    for (auto &pr : g_prof.pairs[kind]) {
        CDR_CUDA(cudaEventSynchronize(pr.second));
        float ms = 0.f;
        CDR_CUDA(cudaEventElapsedTime(&ms, pr.first, pr.second));
        tot += ms;
    }
    if (out_total_ms) *out_total_ms = tot;
    if (out_launches) *out_launches = (int64_t)g_prof.pairs[kind].size();
    return CDR_OK;
}

extern "C" int32_t cdr_prof_read_launches(int32_t kind, double *out_ms, int64_t max_n, int64_t *out_n)
{
    CDR_REQUIRE((kind == 0 || kind == 1) && out_ms && max_n >= 0, CDR_ERR_INVALID, "cdr_prof_read_launches: bad arguments");
    std::lock_guard<std::mutex> lk(g_prof.mu);
    int64_t n = 0;
    for (auto &pr : g_prof.pairs[kind]) {
        if (n >= max_n) break;
        CDR_CUDA(cudaEventSynchronize(pr.second));
        float ms = 0.f;
        CDR_CUDA(cudaEventElapsedTime(&ms, pr.first, pr.second));
        out_ms[n++] = ms;
    }
    if (out_n) *out_n = n;
    return CDR_OK;
}

// ---------------------------------------------------------------------------- exact lane
static int check_search_args(const char *fn, cdr_store *s, const void *q, int nq, int k,
                             const void *o1, const void *o2, const void *o3)
{
    CDR_REQUIRE(s != nullptr, CDR_ERR_INVALID, "%s: store is NULL", fn);
    CDR_REQUIRE(s->finalized, CDR_ERR_STATE, "%s: store not finalized", fn);
    CDR_REQUIRE(nq >= 0 && (nq == 0 || q != nullptr), CDR_ERR_INVALID, "%s: bad query batch", fn);
    CDR_REQUIRE(k >= 1 && k <= CDR_MAX_K, CDR_ERR_UNSUPPORTED, "%s: k=%d outside [1,%d]", fn, k, CDR_MAX_K);
    CDR_REQUIRE(o1 && o2 && o3, CDR_ERR_INVALID, "%s: output buffers required", fn);
    return CDR_OK;
}

static int32_t search_exact_dev(const char *fn, bool share_reads, cdr_store *s, const float *q_dev, int32_t nq, int32_t k,
                                const uint32_t *allow_dev, double *out_score_dev, int64_t *out_id_dev,
                                int32_t *out_n_dev, void *stream, const PeerLink *peer = nullptr)
{
    int rc = check_search_args(fn, s, q_dev, nq, k, out_score_dev, out_id_dev, out_n_dev);
    if (rc != CDR_OK) return rc;
    CDR_REQUIRE(s->emb_f32 != nullptr, CDR_ERR_STATE, "%s: store has no fp32 rows (created without CDR_STORE_FP32)", fn);
    if (nq == 0) return CDR_OK;
    DeviceGuard g(s->device);
    cudaStream_t st = (cudaStream_t)stream;
    std::lock_guard<std::mutex> lk(s->mu);
    ScanWorkspace &ws = s->ws[st];
    // rows with embedding IS NULL are excluded even when the caller passes no filter
    const uint32_t *allow = allow_dev ? allow_dev : (s->any_invalid ? s->valid : nullptr);
    // grid.y is limited to 65535 query groups per launch (a fused exchange serves at most its group's max_nq queries)
    CDR_REQUIRE(peer == nullptr || nq <= peer->max_nq, CDR_ERR_INVALID, "%s: %d queries in one exchange epoch", fn, nq);
    for (int q0 = 0; q0 < nq; q0 += 32768) {
        const int m = (nq - q0) < 32768 ? (nq - q0) : 32768;
        rc = cdr_exact_scan_launch(s, ws, q_dev + (size_t)q0 * s->dim, m, allow, k,
                                   out_score_dev + (size_t)q0 * k, out_id_dev + (size_t)q0 * k,
                                   out_n_dev + q0, st, share_reads, nullptr, peer);
        if (rc != CDR_OK) return rc;
    }
    return CDR_OK;
}

int32_t cdr_search_exact_peer(bool share_reads, cdr_store *s, const float *q_dev, int32_t nq, int32_t k,
                              const uint32_t *allow_dev, double *out_score_dev, int64_t *out_id_dev, int32_t *out_n_dev,
                              void *stream, const PeerLink *peer)
{
    return search_exact_dev("cdr_search_sharded (exact lane)", share_reads, s, q_dev, nq, k, allow_dev, out_score_dev,
                            out_id_dev, out_n_dev, stream, peer);
}

extern "C" int32_t cdr_search_exact_f32(cdr_store *s, const float *q_dev, int32_t nq, int32_t k,
                                        const uint32_t *allow_dev, double *out_score_dev,
                                        int64_t *out_id_dev, int32_t *out_n_dev, void *stream)
{
    return search_exact_dev("cdr_search_exact_f32", false, s, q_dev, nq, k, allow_dev, out_score_dev, out_id_dev,
                            out_n_dev, stream);
}

extern "C" int32_t cdr_search_exact_f32_shared(cdr_store *s, const float *q_dev, int32_t nq, int32_t k,
                                               const uint32_t *allow_dev, double *out_score_dev,
                                               int64_t *out_id_dev, int32_t *out_n_dev, void *stream)
{
    return search_exact_dev("cdr_search_exact_f32_shared", true, s, q_dev, nq, k, allow_dev, out_score_dev,
                            out_id_dev, out_n_dev, stream);
}

extern "C" int32_t cdr_search_scan_bf16(cdr_store *s, const float *q_dev, int32_t nq, int32_t k,
                                        const uint32_t *allow_dev, double *out_score_dev,
                                        int64_t *out_id_dev, int32_t *out_n_dev, void *stream)
{
    return cdr_search_scan_bf16_peer(s, q_dev, nq, k, allow_dev, out_score_dev, out_id_dev, out_n_dev, stream, nullptr);
}

int32_t cdr_search_scan_bf16_peer(cdr_store *s, const float *q_dev, int32_t nq, int32_t k, const uint32_t *allow_dev,
                                  double *out_score_dev, int64_t *out_id_dev, int32_t *out_n_dev, void *stream,
                                  const PeerLink *peer)
{
    const char *fn = "cdr_search_scan_bf16";
    CDR_REQUIRE(peer == nullptr || nq <= peer->max_nq, CDR_ERR_INVALID, "%s: %d queries in one exchange epoch", fn, nq);
    int rc = check_search_args(fn, s, q_dev, nq, k, out_score_dev, out_id_dev, out_n_dev);
    if (rc != CDR_OK) return rc;
    CDR_REQUIRE(s->emb_bf16 != nullptr, CDR_ERR_STATE, "%s: store has no bf16 rows (created without CDR_STORE_BF16)", fn);
    if (nq == 0) return CDR_OK;
    DeviceGuard g(s->device);
    cudaStream_t st = (cudaStream_t)stream;
    std::lock_guard<std::mutex> lk(s->mu);
    ScanWorkspace &ws = s->ws[st];
    const uint32_t *allow = allow_dev ? allow_dev : (s->any_invalid ? s->valid : nullptr);
    for (int q0 = 0; q0 < nq; q0 += 32768) {
        const int m = (nq - q0) < 32768 ? (nq - q0) : 32768;
        rc = cdr_bf16_scan_launch(s, ws, q_dev + (size_t)q0 * s->dim, m, allow, k, out_score_dev + (size_t)q0 * k,
                                  out_id_dev + (size_t)q0 * k, out_n_dev + q0, st, nullptr, nullptr, peer);
        if (rc != CDR_OK) return rc;
    }
    return CDR_OK;
}

// Shared host-buffer wrapper: H2D queries, run `fn`, D2H results, synchronise.
typedef int32_t (*search_fn)(cdr_store *, const float *, int32_t, int32_t, const uint32_t *, double *,
                             int64_t *, int32_t *, void *);

static int32_t search_host(const char *name, search_fn fn, cdr_store *s, const float *q_host,
                           int32_t nq, int32_t k, const uint32_t *allow_dev, double *out_score_host,
                           int64_t *out_id_host, int32_t *out_n_host, void *stream)
{
    int rc = check_search_args(name, s, q_host, nq, k, out_score_host, out_id_host, out_n_host);
    if (rc != CDR_OK) return rc;
    if (nq == 0) return CDR_OK;
    DeviceGuard g(s->device);
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char *hp = nullptr;
    const size_t qb = (size_t)nq * s->dim * 4;
    const size_t sb = (size_t)nq * k * 8, ib = (size_t)nq * k * 8, nb = (size_t)nq * 4;
    const size_t qb_al = (qb + 255) & ~(size_t)255;
    // this thread's device staging: [queries | scores | ids | counts]
    unsigned char *dstage = (unsigned char *)cdr_thread_device(s->device, qb_al + sb + ib + nb + 256);
    if (!dstage) return CDR_ERR_OOM;
    float *qd = (float *)dstage;
    unsigned char *od = dstage + qb_al;
    // small requests (the per-request case: one query, 4 KB) go through this thread's pinned mirror: one
    // H2D and one D2H DMA per call instead of driver-staged copies from / to pageable caller memory; large
    // batches are copied straight from / to the caller's buffers (a pinned caller buffer then DMAs without
    // a host memcpy)
    const bool staged = qb <= ((size_t)256 << 10);
    if (staged) {
        hp = (unsigned char *)cdr_thread_pinned(qb_al + sb + ib + nb);
        if (!hp) return CDR_ERR_OOM;
    }
    double *sd = (double *)od;
    int64_t *idd = (int64_t *)(od + sb);
    int32_t *nd = (int32_t *)(od + sb + ib);
    if (staged) memcpy(hp, q_host, qb);
    CDR_CUDA(cudaMemcpyAsync(qd, staged ? (const void *)hp : (const void *)q_host, qb, cudaMemcpyHostToDevice, st));
    rc = fn(s, qd, nq, k, allow_dev, sd, idd, nd, stream);
    if (rc != CDR_OK) return rc;
    if (staged) {
        CDR_CUDA(cudaMemcpyAsync(hp + qb_al, od, sb + ib + nb, cudaMemcpyDeviceToHost, st));
        CDR_CUDA(cudaStreamSynchronize(st));
        memcpy(out_score_host, hp + qb_al, sb);
        memcpy(out_id_host, hp + qb_al + sb, ib);
        memcpy(out_n_host, hp + qb_al + sb + ib, nb);
        return CDR_OK;
    }
    CDR_CUDA(cudaMemcpyAsync(out_score_host, sd, sb, cudaMemcpyDeviceToHost, st));
    CDR_CUDA(cudaMemcpyAsync(out_id_host, idd, ib, cudaMemcpyDeviceToHost, st));
    CDR_CUDA(cudaMemcpyAsync(out_n_host, nd, nb, cudaMemcpyDeviceToHost, st));
    CDR_CUDA(cudaStreamSynchronize(st));
    return CDR_OK;
}

extern "C" int32_t cdr_search_exact_f32_host(cdr_store *s, const float *q_host, int32_t nq, int32_t k,
                                             const uint32_t *allow_dev, double *out_score_host,
                                             int64_t *out_id_host, int32_t *out_n_host, void *stream)
{
    return search_host("cdr_search_exact_f32_host", cdr_search_exact_f32, s, q_host, nq, k, allow_dev,
                       out_score_host, out_id_host, out_n_host, stream);
}

extern "C" int32_t cdr_search_exact_f32_shared_host(cdr_store *s, const float *q_host, int32_t nq, int32_t k,
                                                    const uint32_t *allow_dev, double *out_score_host,
                                                    int64_t *out_id_host, int32_t *out_n_host, void *stream)
{
    return search_host("cdr_search_exact_f32_shared_host", cdr_search_exact_f32_shared, s, q_host, nq, k, allow_dev,
                       out_score_host, out_id_host, out_n_host, stream);
}

extern "C" int32_t cdr_search_scan_bf16_host(cdr_store *s, const float *q_host, int32_t nq, int32_t k,
                                             const uint32_t *allow_dev, double *out_score_host,
                                             int64_t *out_id_host, int32_t *out_n_host, void *stream)
{
    return search_host("cdr_search_scan_bf16_host", cdr_search_scan_bf16, s, q_host, nq, k, allow_dev,
                       out_score_host, out_id_host, out_n_host, stream);
}

extern "C" int32_t cdr_search_batch_bf16_host(cdr_store *s, const float *q_host, int32_t nq, int32_t k,
                                              const uint32_t *allow_dev, double *out_score_host,
                                              int64_t *out_id_host, int32_t *out_n_host, void *stream)
{
    return search_host("cdr_search_batch_bf16_host", cdr_search_batch_bf16, s, q_host, nq, k, allow_dev,
                       out_score_host, out_id_host, out_n_host, stream);
}
