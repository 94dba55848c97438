// rrf.cu -- K5: reciprocal-rank fusion, bit-exact with the reference.
//
// Replaces app/retrieve.py:245-260 (_rrf_merge):
//     for lane in lanes (dict order: bm25, tech_tokens, dense -- app/retrieve.py:537-547):
//         for rank, row in enumerate(rows, start=1):
//             scores[key] = scores.get(key, 0.0) + 1.0 / (k + rank)
//     sorted(scores.items(), key=score, reverse=True)      # stable: ties keep first-seen order
// One CTA per query.  Every distinct id is owned by the thread of its first occurrence, which
// walks the remaining items IN SEQUENCE ORDER and accumulates with IEEE fp64 add/div (explicit
// _rn intrinsics, no FMA contraction), so the association order equals Python's.  The final
// order is produced by rank counting with the key (score desc, first-seen index asc), which is
// exactly what a stable descending sort yields.
#include "common.cuh"

namespace {

constexpr int kRrfThreads = 256;

struct RrfParams {
    const int64_t *lane_ids;
    const int32_t *lane_offsets;   // [nq*L + 1]
    int L;
    int rrf_k;
    int max_out;
    int64_t *out_ids;
    double *out_scores;
    uint32_t *out_mask;
    int32_t *out_n;
};

__global__ void __launch_bounds__(kRrfThreads) rrf_merge_kernel(const RrfParams p)
{
    __shared__ int64_t s_id[CDR_RRF_MAX_ITEMS];
    __shared__ double s_score[CDR_RRF_MAX_ITEMS];     // score of the id owned by item i (first occurrence)
    __shared__ uint16_t s_rank[CDR_RRF_MAX_ITEMS];    // 1-based rank inside its lane
    __shared__ uint8_t s_lane[CDR_RRF_MAX_ITEMS];
    __shared__ uint8_t s_first[CDR_RRF_MAX_ITEMS];    // 1 when item i is the first occurrence of its id
    __shared__ double s_term[CDR_RRF_MAX_ITEMS];      // 1 / (k + rank) of item i
    __shared__ uint64_t s_key[CDR_RRF_MAX_ITEMS];     // order key of s_score for the owners
    __shared__ uint32_t s_mask[CDR_RRF_MAX_ITEMS];
    __shared__ int s_total, s_unique;

    const int q = blockIdx.x;
    const int32_t *off = p.lane_offsets + (size_t)q * p.L;
    const int begin = off[0];
    int total = off[p.L] - begin;
    if (total > CDR_RRF_MAX_ITEMS) total = CDR_RRF_MAX_ITEMS;   // host validates; defensive clamp
    if (threadIdx.x == 0) { s_total = total; s_unique = 0; }

    // load items in sequence order (lane-major, rank-minor)
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int g = begin + i;
        int l = 0;
        while (l + 1 < p.L && g >= off[l + 1]) ++l;
        s_id[i] = p.lane_ids[g];
        s_lane[i] = (uint8_t)l;
        s_rank[i] = (uint16_t)(g - off[l] + 1);
        s_term[i] = __ddiv_rn(1.0, (double)(p.rrf_k + (g - off[l] + 1)));
    }
    __syncthreads();

    // One branch-free pass per item over the whole sequence (data-dependent branches in these short loops cost more than
    // the arithmetic they skip): is an earlier item the same id?  If not, item i owns the id and adds the terms of its
    // occurrences IN SEQUENCE ORDER -- every other step adds +0.0, which leaves the sum's bits alone -- and collects the
    // lanes that carry it.
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int64_t id = s_id[i];
        int earlier = 0;
        double sc = 0.0;
        uint32_t mask = 0;
#pragma unroll 8
        for (int j = 0; j < total; ++j) {
            const bool eq = s_id[j] == id;
            earlier |= (int)(eq & (j < i));
            const bool mine = eq & (j >= i);
            sc = __dadd_rn(sc, mine ? s_term[j] : 0.0);
            mask |= mine ? (1u << s_lane[j]) : 0u;
        }
        s_first[i] = earlier ? 0 : 1;
        s_score[i] = sc;
        s_key[i] = earlier ? 0ull : cdr_order_f64(sc);          // compared as integers below (same order as the doubles)
        s_mask[i] = mask;
    }
    __syncthreads();

    int uniq_local = 0;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        if (!s_first[i]) continue;
        ++uniq_local;
        const uint64_t key = s_key[i];
        int pos = 0;
#pragma unroll 8
        for (int j = 0; j < total; ++j) {
            const uint64_t o = s_key[j];
            pos += (int)(s_first[j] != 0) & (int)((o > key) | ((o == key) & (j < i)));
        }
        if (pos < p.max_out) {
            const size_t o = (size_t)q * p.max_out + pos;
            p.out_ids[o] = s_id[i];
            p.out_scores[o] = s_score[i];
            p.out_mask[o] = s_mask[i];
        }
    }
    if (uniq_local) atomicAdd(&s_unique, uniq_local);
    __syncthreads();
    const int n_out = s_unique < p.max_out ? s_unique : p.max_out;
    for (int i = n_out + threadIdx.x; i < p.max_out; i += blockDim.x) {
        const size_t o = (size_t)q * p.max_out + i;
        p.out_ids[o] = -1;
        p.out_scores[o] = 0.0;
        p.out_mask[o] = 0;
    }
    if (threadIdx.x == 0) p.out_n[q] = n_out;
}

}  // namespace

extern "C" int32_t cdr_rrf_merge(const int64_t *lane_ids_dev, const int32_t *lane_offsets_dev,
                                 int32_t nq, int32_t L, int32_t rrf_k, int32_t max_out,
                                 int64_t *out_ids_dev, double *out_scores_dev,
                                 uint32_t *out_lane_mask_dev, int32_t *out_n_dev, void *stream)
{
    CDR_REQUIRE(nq >= 0 && L >= 1 && L <= 32 && max_out >= 1, CDR_ERR_INVALID,
                "cdr_rrf_merge: need nq >= 0, 1 <= L <= 32, max_out >= 1");
    CDR_REQUIRE(lane_offsets_dev && out_ids_dev && out_scores_dev && out_lane_mask_dev && out_n_dev,
                CDR_ERR_INVALID, "cdr_rrf_merge: NULL argument");
    if (nq == 0) return CDR_OK;
    RrfParams p;
    p.lane_ids = lane_ids_dev;
    p.lane_offsets = lane_offsets_dev;
    p.L = L;
    p.rrf_k = rrf_k;
    p.max_out = max_out;
    p.out_ids = out_ids_dev;
    p.out_scores = out_scores_dev;
    p.out_mask = out_lane_mask_dev;
    p.out_n = out_n_dev;
    rrf_merge_kernel<<<nq, kRrfThreads, 0, (cudaStream_t)stream>>>(p);
    CDR_LAUNCH_CHECK();
    return CDR_OK;
}

extern "C" int32_t cdr_rrf_merge_host(const int64_t *lane_ids_host, const int32_t *lane_offsets_host,
                                      int32_t nq, int32_t L, int32_t rrf_k, int32_t max_out,
                                      int64_t *out_ids_host, double *out_scores_host,
                                      uint32_t *out_lane_mask_host, int32_t *out_n_host, void *stream)
{
    CDR_REQUIRE(nq >= 0 && L >= 1 && L <= 32 && max_out >= 1 && lane_offsets_host, CDR_ERR_INVALID,
                "cdr_rrf_merge_host: bad arguments");
    if (nq == 0) return CDR_OK;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        cdr_set_error("cdr_rrf_merge_host: no CUDA device visible; this engine has no CPU fallback");
        return CDR_ERR_NO_DEVICE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t total = lane_offsets_host[(size_t)nq * L];
    for (int q = 0; q < nq; ++q) {
        const int64_t t = lane_offsets_host[(size_t)(q + 1) * L] - lane_offsets_host[(size_t)q * L];
        CDR_REQUIRE(t >= 0 && t <= CDR_RRF_MAX_ITEMS, CDR_ERR_INVALID,
                    "cdr_rrf_merge_host: query %d has %lld lane items (max %d)", q, (long long)t,
                    CDR_RRF_MAX_ITEMS);
    }
    const size_t b_ids = (size_t)(total > 0 ? total : 1) * 8, b_off = ((size_t)nq * L + 1) * 4;
    const size_t b_oid = (size_t)nq * max_out * 8, b_osc = (size_t)nq * max_out * 8;
    const size_t b_om = (size_t)nq * max_out * 4, b_on = (size_t)nq * 4;
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t tot = up(b_ids) + up(b_off) + up(b_oid) + up(b_osc) + up(b_om) + up(b_on);
    int cur_dev = 0;
    CDR_CUDA(cudaGetDevice(&cur_dev));
    unsigned char *buf = (unsigned char *)cdr_thread_device(cur_dev, tot);      // this thread's device staging
    if (!buf) return CDR_ERR_OOM;
    unsigned char *c = buf;
    int64_t *d_ids = (int64_t *)c; c += up(b_ids);
    int32_t *d_off = (int32_t *)c; c += up(b_off);
    int64_t *d_oid = (int64_t *)c; c += up(b_oid);
    double *d_osc = (double *)c; c += up(b_osc);
    uint32_t *d_om = (uint32_t *)c; c += up(b_om);
    int32_t *d_on = (int32_t *)c;
    if (total > 0) CDR_CUDA(cudaMemcpyAsync(d_ids, lane_ids_host, (size_t)total * 8, cudaMemcpyHostToDevice, st));
    CDR_CUDA(cudaMemcpyAsync(d_off, lane_offsets_host, b_off, cudaMemcpyHostToDevice, st));
    int rc = cdr_rrf_merge(d_ids, d_off, nq, L, rrf_k, max_out, d_oid, d_osc, d_om, d_on, stream);
    if (rc != CDR_OK) return rc;
    CDR_CUDA(cudaMemcpyAsync(out_ids_host, d_oid, b_oid, cudaMemcpyDeviceToHost, st));
    CDR_CUDA(cudaMemcpyAsync(out_scores_host, d_osc, b_osc, cudaMemcpyDeviceToHost, st));
    CDR_CUDA(cudaMemcpyAsync(out_lane_mask_host, d_om, b_om, cudaMemcpyDeviceToHost, st));
    CDR_CUDA(cudaMemcpyAsync(out_n_host, d_on, b_on, cudaMemcpyDeviceToHost, st));
    CDR_CUDA(cudaStreamSynchronize(st));
    return CDR_OK;
}
