// topk_merge.cu -- K4: merge R per-shard top-k lists per query into the global top-k.
//
// Multi-GPU step of the row-sharded corpus: after the NCCL all-gather of each
// rank's [nq,k] (score f64, id i64) lists, every rank runs this deterministic merge, so all
// ranks hold the identical result.  Ordering rule as everywhere: score desc, NaN last, id asc.
// One CTA per query; R*k <= 4096 entries ranked by counting in shared memory.
#include "common.cuh"

#include <mutex>

namespace {

constexpr int kMergeMax = 4096;

struct MergeParams {
    const double *scores;   // [R, nq, k]
    const int64_t *ids;     // [R, nq, k]
    const int32_t *n;       // [R, nq]
    int R, nq, k;
    double *out_score;
    int64_t *out_id;
    int32_t *out_n;
};

__global__ void __launch_bounds__(256) topk_merge_kernel(const MergeParams p)
{
    extern __shared__ unsigned char raw[];
    uint64_t *s_key = reinterpret_cast<uint64_t *>(raw);                   // order keys of the scores (integer compares)
    int64_t *s_id = reinterpret_cast<int64_t *>(s_key + (size_t)p.R * p.k);
    __shared__ int s_cnt;
    const int q = blockIdx.x;
    const int tot = p.R * p.k;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    int *s_n = reinterpret_cast<int *>(s_id + (size_t)p.R * p.k);          // valid entries of list r
    for (int r = threadIdx.x; r < p.R; r += blockDim.x) {
        int n = p.n[(size_t)r * p.nq + q];
        n = n < 0 ? 0 : (n > p.k ? p.k : n);
        s_n[r] = n;
        if (n) atomicAdd(&s_cnt, n);
    }
    __syncthreads();
    for (int e = threadIdx.x; e < tot; e += blockDim.x) {
        const int r = e / p.k, i = e - r * p.k;
        const bool ok = i < s_n[r];
        const size_t g = ((size_t)r * p.nq + q) * p.k + i;
        s_key[e] = ok ? cdr_order_f64(p.scores[g]) : 0ull;
        s_id[e] = ok ? p.ids[g] : INT64_MAX;
    }
    __syncthreads();
    // every list is ordered: global rank of entry i of list r = i + (entries of the other lists before it), each found by
    // a branch-free binary search
    for (int e = threadIdx.x; e < tot; e += blockDim.x) {
        const int r = e / p.k, i = e - r * p.k;
        if (i >= s_n[r]) continue;
        const uint64_t key = s_key[e];
        const int64_t id = s_id[e];
        int rank = i;
        for (int o = 0; o < p.R; ++o)
            if (o != r) rank += cdr_sorted_count_before(s_key + o * p.k, s_id + o * p.k, s_n[o], p.k, key, id);
        if (rank < p.k) {
            p.out_score[(size_t)q * p.k + rank] = p.scores[((size_t)r * p.nq + q) * p.k + i];
            p.out_id[(size_t)q * p.k + rank] = id;
        }
    }
    const int n_out = s_cnt < p.k ? s_cnt : p.k;
    for (int i = n_out + threadIdx.x; i < p.k; i += blockDim.x) {
        p.out_score[(size_t)q * p.k + i] = __longlong_as_double(0x7FF8000000000000ll);
        p.out_id[(size_t)q * p.k + i] = -1;
    }
    if (threadIdx.x == 0) p.out_n[q] = n_out;
}

}  // namespace

extern "C" int32_t cdr_topk_merge(const double *scores_dev, const int64_t *ids_dev,
                                  const int32_t *n_dev, int32_t R, int32_t nq, int32_t k,
                                  double *out_score_dev, int64_t *out_id_dev, int32_t *out_n_dev,
                                  void *stream)
{
    CDR_REQUIRE(scores_dev && ids_dev && n_dev && out_score_dev && out_id_dev && out_n_dev,
                CDR_ERR_INVALID, "cdr_topk_merge: NULL argument");
    CDR_REQUIRE(R >= 1 && nq >= 0 && k >= 1 && (int64_t)R * k <= kMergeMax, CDR_ERR_INVALID,
                "cdr_topk_merge: need R >= 1, k >= 1, R*k <= %d (got R=%d k=%d)", kMergeMax, R, k);
    if (nq == 0) return CDR_OK;
    const size_t smem = (size_t)R * k * 16 + (size_t)R * 4;
    {
        // the opt-in is per device: one flag per device of the calling thread, set under a lock
        static std::mutex attr_mu;
        static bool attr_set[64] = {false};
        int dev = 0;
        CDR_CUDA(cudaGetDevice(&dev));
        std::lock_guard<std::mutex> attr_lock(attr_mu);
        if (!attr_set[dev & 63] && smem > 48 * 1024) {
            CDR_CUDA(cudaFuncSetAttribute(topk_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          kMergeMax * 20));
            attr_set[dev & 63] = true;
        }
    }
    MergeParams p{scores_dev, ids_dev, n_dev, R, nq, k, out_score_dev, out_id_dev, out_n_dev};
    topk_merge_kernel<<<nq, 256, smem, (cudaStream_t)stream>>>(p);
    CDR_LAUNCH_CHECK();
    return CDR_OK;
}
