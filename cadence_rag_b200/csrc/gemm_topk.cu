// gemm_topk.cu -- K2 placeholder (replaced by the tcgen05 kernel in the next commit).
#include "common.cuh"

extern "C" int32_t cdr_search_batch_bf16(cdr_store *s, const float *q_dev, int32_t nq, int32_t k,
                                         const uint32_t *allow_dev, double *out_score_dev,
                                         int64_t *out_id_dev, int32_t *out_n_dev, void *stream)
{
    cdr_set_error("cdr_search_batch_bf16: not built yet");
    return CDR_ERR_UNSUPPORTED;
}
