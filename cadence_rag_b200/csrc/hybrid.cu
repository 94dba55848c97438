// hybrid.cu -- one fused call per /retrieve request batch and table:
//   K6 filter bitmap + COUNT(*)  ->  K1 exact dense lane (+ fp64 re-score)  ->  tech_tokens lane
//   ->  lane assembly  ->  K5 RRF,  with ONE host->device copy, ONE device->host copy and ONE
//   stream synchronisation.
//
// Replaces, for one table, the sequence retrieve_evidence runs per request (app/retrieve.py:445-550):
//   _fetch_chunks_tech (:183-209)            -> tech_tokens lane
//   _estimate_dense_candidates (:303-323)    -> COUNT(*) (returned; the planner stays in Python)
//   _fetch_chunks_dense (:326-354)           -> dense lane
//   _rrf_merge({"bm25","tech_tokens","dense"}) (:245-260, :537-550) -> fused ranks (bit-exact)
// The BM25 lane (pg_search, out of scope) is an input: its ranked ids are passed in.
#include "common.cuh"

#include <cstring>

struct cdr_tech_index;
int cdr_filter_launch(cdr_store *s, const uint32_t *bm_dev, int64_t n_call_slots, int has_from, int64_t date_from_us,
                      int has_to, int64_t date_to_us, int has_tags, uint64_t tag_any, uint32_t *out_allow_dev,
                      unsigned long long *cnt_dev, cudaStream_t st);
int cdr_tech_lane_launch(cdr_tech_index *ix, const int32_t *d_tok, const int32_t *d_nt, int nq, int max_tokens,
                         const uint32_t *d_bm, int64_t n_call_slots, int has_date_from, int64_t date_from_us,
                         int has_date_to, int64_t date_to_us, int has_tag_filter, uint64_t tag_any, int limit,
                         int64_t *d_oid, int32_t *d_on, cudaStream_t st);
cdr_store *cdr_tech_index_store(cdr_tech_index *ix);

namespace {

struct AssembleParams {
    const int32_t *bm25_off;   // [nq+1] or nullptr
    const int64_t *bm25_ids;
    const int64_t *tech_ids;   // [nq, tech_limit]
    const int32_t *tech_n;     // [nq]
    int tech_limit;
    const int64_t *dense_ids;  // [nq, dense_k] or nullptr (dense lane disabled => L = 2)
    const int32_t *dense_n;
    int dense_k;
    int L;
    int64_t *lane_ids;         // compacted: query-major, lane-major, rank-minor
    int32_t *lane_off;         // [nq*L + 1]
    int nq;
};

__device__ __forceinline__ int lane_len(const AssembleParams &p, int q, int l)
{
    if (l == 0) return p.bm25_off ? p.bm25_off[q + 1] - p.bm25_off[q] : 0;
    if (l == 1) return p.tech_n[q];
    return p.dense_n[q];
}

// One CTA per query: begin = sum of the item counts of all earlier queries (nq is small), then
// copy the three lanes back to back and write the L (+1 for the last query) offsets.
__global__ void __launch_bounds__(128) assemble_lanes_kernel(const AssembleParams p)
{
    __shared__ int s_part[4];
    const int q = blockIdx.x;
    int acc = 0;
    for (int o = threadIdx.x; o < q; o += blockDim.x)
        for (int l = 0; l < p.L; ++l) acc += lane_len(p, o, l);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    int pos = s_part[0] + s_part[1] + s_part[2] + s_part[3];
    for (int l = 0; l < p.L; ++l) {
        const int n = lane_len(p, q, l);
        const int64_t *src = l == 0 ? (p.bm25_off ? p.bm25_ids + p.bm25_off[q] : nullptr)
                             : l == 1 ? p.tech_ids + (size_t)q * p.tech_limit
                                      : p.dense_ids + (size_t)q * p.dense_k;
        if (threadIdx.x == 0) p.lane_off[(size_t)q * p.L + l] = pos;
        for (int i = threadIdx.x; i < n; i += blockDim.x) p.lane_ids[pos + i] = src[i];
        pos += n;
    }
    if (q == p.nq - 1 && threadIdx.x == 0) p.lane_off[(size_t)p.nq * p.L] = pos;
}

inline size_t up16(size_t x) { return (x + 15) & ~(size_t)15; }

}  // namespace

extern "C" int32_t cdr_hybrid_retrieve_host(
    cdr_store *s, cdr_tech_index *tech_index, const cdr_filter_spec *filter, const float *q_host, int32_t nq,
    int32_t dense_k, const int32_t *token_ids_host, const int32_t *n_tokens_host, int32_t max_tokens,
    int32_t tech_limit, const int64_t *bm25_ids_host, const int32_t *bm25_offsets_host, int32_t rrf_k,
    int32_t max_out, int64_t *out_count_host, int64_t *out_dense_ids_host, double *out_dense_scores_host,
    int32_t *out_dense_n_host, int64_t *out_tech_ids_host, int32_t *out_tech_n_host, int64_t *out_fused_ids_host,
    double *out_fused_scores_host, uint32_t *out_fused_mask_host, int32_t *out_fused_n_host, void *stream)
{
    const char *fn = "cdr_hybrid_retrieve_host";
    CDR_REQUIRE(s != nullptr, CDR_ERR_INVALID, "%s: store is NULL", fn);
    CDR_REQUIRE(s->finalized, CDR_ERR_STATE, "%s: store not finalized", fn);
    CDR_REQUIRE(nq >= 0 && nq <= 4096, CDR_ERR_INVALID, "%s: nq=%d outside [0,4096]", fn, nq);
    const bool dense = q_host != nullptr;
    if (dense) {
        CDR_REQUIRE(s->emb_f32 != nullptr, CDR_ERR_STATE, "%s: the dense lane of the fused call needs fp32 rows", fn);
        CDR_REQUIRE(dense_k >= 1 && dense_k <= CDR_MAX_K, CDR_ERR_UNSUPPORTED, "%s: dense_k=%d outside [1,%d]", fn,
                    dense_k, CDR_MAX_K);
        CDR_REQUIRE(out_dense_ids_host && out_dense_scores_host && out_dense_n_host, CDR_ERR_INVALID,
                    "%s: dense outputs required", fn);
    } else {
        dense_k = 0;
    }
    const bool tech = tech_index != nullptr && token_ids_host != nullptr && n_tokens_host != nullptr;
    CDR_REQUIRE(tech_limit >= 1 && tech_limit <= CDR_MAX_K, CDR_ERR_UNSUPPORTED, "%s: tech_limit=%d outside [1,%d]", fn,
                tech_limit, CDR_MAX_K);
    CDR_REQUIRE(!tech || (max_tokens >= 1 && max_tokens <= 32), CDR_ERR_INVALID, "%s: need 1 <= max_tokens <= 32", fn);
    CDR_REQUIRE(!tech || cdr_tech_index_store(tech_index) == s, CDR_ERR_INVALID, "%s: tech index belongs to another store", fn);
    CDR_REQUIRE(out_tech_ids_host && out_tech_n_host && out_fused_ids_host && out_fused_scores_host &&
                out_fused_mask_host && out_fused_n_host, CDR_ERR_INVALID, "%s: output buffers required", fn);
    CDR_REQUIRE(max_out >= 1 && rrf_k >= 0, CDR_ERR_INVALID, "%s: need max_out >= 1, rrf_k >= 0", fn);
    CDR_REQUIRE((bm25_ids_host == nullptr) == (bm25_offsets_host == nullptr) || bm25_offsets_host != nullptr,
                CDR_ERR_INVALID, "%s: bm25 ids without offsets", fn);
    if (nq == 0) return CDR_OK;
    int64_t bm25_total = 0, bm25_max = 0;
    if (bm25_offsets_host) {
        CDR_REQUIRE(bm25_offsets_host[0] == 0, CDR_ERR_INVALID, "%s: bm25 offsets must start at 0", fn);
        for (int q = 0; q < nq; ++q) {
            const int64_t n = (int64_t)bm25_offsets_host[q + 1] - bm25_offsets_host[q];
            CDR_REQUIRE(n >= 0, CDR_ERR_INVALID, "%s: bm25 offsets not monotone at query %d", fn, q);
            if (n > bm25_max) bm25_max = n;
        }
        bm25_total = bm25_offsets_host[nq];
        CDR_REQUIRE(bm25_total == 0 || bm25_ids_host, CDR_ERR_INVALID, "%s: bm25 ids missing", fn);
    }
    CDR_REQUIRE(bm25_max + tech_limit + dense_k <= CDR_RRF_MAX_ITEMS, CDR_ERR_INVALID,
                "%s: %lld lane items per query exceed %d", fn, (long long)(bm25_max + tech_limit + dense_k),
                CDR_RRF_MAX_ITEMS);
    const bool has_filter = filter != nullptr && (filter->call_slot_bitmap_host || filter->has_date_from ||
                                                  filter->has_date_to || filter->has_tag_filter);
    const bool has_bm = has_filter && filter->call_slot_bitmap_host != nullptr;
    CDR_REQUIRE(!has_bm || filter->n_call_slots >= 0, CDR_ERR_INVALID, "%s: n_call_slots < 0", fn);

    DeviceGuard g(s->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int dim = s->dim;
    const int L = dense ? 3 : 2;
    const size_t bm_words = has_bm ? (size_t)((filter->n_call_slots + 31) / 32) : 0;

    // ---- packed request (host -> device) and packed response (device -> host)
    size_t in_q = 0, in_tok, in_nt, in_boff, in_bids, in_bm, in_end;
    in_tok = up16(in_q + (dense ? (size_t)nq * dim * 4 : 0));
    in_nt = up16(in_tok + (tech ? (size_t)nq * max_tokens * 4 : 0));
    in_boff = up16(in_nt + (tech ? (size_t)nq * 4 : 0));
    in_bids = up16(in_boff + (bm25_offsets_host ? (size_t)(nq + 1) * 4 : 0));
    in_bm = up16(in_bids + (size_t)bm25_total * 8);
    in_end = up16(in_bm + (bm_words + 1) * 4);
    size_t o_cnt = in_end, o_did, o_dsc, o_dn, o_tid, o_tn, o_fid, o_fsc, o_fm, o_fn, o_end;
    o_did = up16(o_cnt + 8);
    o_dsc = up16(o_did + (size_t)nq * dense_k * 8);
    o_dn = up16(o_dsc + (size_t)nq * dense_k * 8);
    o_tid = up16(o_dn + (size_t)nq * 4);
    o_tn = up16(o_tid + (size_t)nq * tech_limit * 8);
    o_fid = up16(o_tn + (size_t)nq * 4);
    o_fsc = up16(o_fid + (size_t)nq * max_out * 8);
    o_fm = up16(o_fsc + (size_t)nq * max_out * 8);
    o_fn = up16(o_fm + (size_t)nq * max_out * 4);
    o_end = up16(o_fn + (size_t)nq * 4);
    // device-only scratch: allow bitmap, compacted lanes, offsets
    const size_t words = (size_t)((s->n_rows + 31) / 32);
    const size_t x_allow = o_end;
    const size_t x_lids = up16(x_allow + (has_filter ? (words + 1) * 4 : 0));
    const size_t x_loff = up16(x_lids + ((size_t)bm25_total + (size_t)nq * (tech_limit + dense_k) + 1) * 8);
    const size_t x_end = up16(x_loff + ((size_t)nq * L + 1) * 4);

    // request / response mirrors in this thread's pinned + device staging.  The store lock is held while the
    // work is ENQUEUED only (the kernels' workspaces are per stream).
    unsigned char *h = (unsigned char *)cdr_thread_pinned(o_end);
    unsigned char *d = (unsigned char *)cdr_thread_device(s->device, x_end);
    if (!h || !d) return CDR_ERR_OOM;
    std::unique_lock<std::mutex> lk(s->mu);
    ScanWorkspace &ws = s->ws[st];

    if (dense) memcpy(h + in_q, q_host, (size_t)nq * dim * 4);
    if (tech) {
        memcpy(h + in_tok, token_ids_host, (size_t)nq * max_tokens * 4);
        memcpy(h + in_nt, n_tokens_host, (size_t)nq * 4);
    }
    if (bm25_offsets_host) {
        memcpy(h + in_boff, bm25_offsets_host, (size_t)(nq + 1) * 4);
        if (bm25_total) memcpy(h + in_bids, bm25_ids_host, (size_t)bm25_total * 8);
    }
    memset(h + in_bm, 0, (bm_words + 1) * 4);
    if (has_bm && bm_words) memcpy(h + in_bm, filter->call_slot_bitmap_host, bm_words * 4);
    memset(h + o_cnt, 0, 8);                                      // zeroes the device counter through the same copy
    CDR_CUDA(cudaMemcpyAsync(d, h, o_cnt + 8, cudaMemcpyHostToDevice, st));

    const uint32_t *d_bm = has_bm ? (const uint32_t *)(d + in_bm) : nullptr;
    const int64_t n_slots = has_bm ? filter->n_call_slots : 0;
    const int f_from = has_filter ? filter->has_date_from : 0, f_to = has_filter ? filter->has_date_to : 0;
    const int f_tags = has_filter ? filter->has_tag_filter : 0;
    const int64_t t_from = has_filter ? filter->date_from_us : 0, t_to = has_filter ? filter->date_to_us : 0;
    const uint64_t tag_any = has_filter ? filter->tag_any : 0;
    int rc;

    // tech_tokens lane first: short, and independent of the dense lane
    int64_t *d_tid = (int64_t *)(d + o_tid);
    int32_t *d_tn = (int32_t *)(d + o_tn);
    if (tech) {
        rc = cdr_tech_lane_launch(tech_index, (const int32_t *)(d + in_tok), (const int32_t *)(d + in_nt), nq,
                                  max_tokens, d_bm, n_slots, f_from, t_from, f_to, t_to, f_tags, tag_any, tech_limit,
                                  d_tid, d_tn, st);
        if (rc != CDR_OK) return rc;
    } else {
        CDR_CUDA(cudaMemsetAsync(d_tid, 0xFF, (size_t)nq * tech_limit * 8, st));
        CDR_CUDA(cudaMemsetAsync(d_tn, 0, (size_t)nq * 4, st));
    }

    // dense lane: WHERE <filters> AND embedding IS NOT NULL, ORDER BY embedding <=> q LIMIT k
    const uint32_t *allow = s->any_invalid ? s->valid : nullptr;
    if (has_filter) {
        uint32_t *d_allow = (uint32_t *)(d + x_allow);
        rc = cdr_filter_launch(s, d_bm, n_slots, f_from, t_from, f_to, t_to, f_tags, tag_any, d_allow,
                               (unsigned long long *)(d + o_cnt), st);
        if (rc != CDR_OK) return rc;
        allow = d_allow;
    }
    if (dense) {
        rc = cdr_exact_scan_launch(s, ws, (const float *)(d + in_q), nq, allow, dense_k, (double *)(d + o_dsc),
                                   (int64_t *)(d + o_did), (int32_t *)(d + o_dn), st, /*share_reads=*/true);
        if (rc != CDR_OK) return rc;
    }

    // lanes -> RRF
    AssembleParams ap;
    ap.bm25_off = bm25_offsets_host ? (const int32_t *)(d + in_boff) : nullptr;
    ap.bm25_ids = (const int64_t *)(d + in_bids);
    ap.tech_ids = d_tid;
    ap.tech_n = d_tn;
    ap.tech_limit = tech_limit;
    ap.dense_ids = dense ? (const int64_t *)(d + o_did) : nullptr;
    ap.dense_n = dense ? (const int32_t *)(d + o_dn) : nullptr;
    ap.dense_k = dense_k;
    ap.L = L;
    ap.lane_ids = (int64_t *)(d + x_lids);
    ap.lane_off = (int32_t *)(d + x_loff);
    ap.nq = nq;
    assemble_lanes_kernel<<<nq, 128, 0, st>>>(ap);
    CDR_LAUNCH_CHECK();
    rc = cdr_rrf_merge(ap.lane_ids, ap.lane_off, nq, L, rrf_k, max_out, (int64_t *)(d + o_fid), (double *)(d + o_fsc),
                       (uint32_t *)(d + o_fm), (int32_t *)(d + o_fn), stream);
    if (rc != CDR_OK) return rc;

    CDR_CUDA(cudaMemcpyAsync(h + o_cnt, d + o_cnt, o_end - o_cnt, cudaMemcpyDeviceToHost, st));
    const int64_t n_valid_now = s->n_valid;
    lk.unlock();
    CDR_CUDA(cudaStreamSynchronize(st));

    if (out_count_host) {
        unsigned long long c;
        memcpy(&c, h + o_cnt, 8);
        *out_count_host = has_filter ? (int64_t)c : n_valid_now;
    }
    if (dense) {
        memcpy(out_dense_ids_host, h + o_did, (size_t)nq * dense_k * 8);
        memcpy(out_dense_scores_host, h + o_dsc, (size_t)nq * dense_k * 8);
        memcpy(out_dense_n_host, h + o_dn, (size_t)nq * 4);
    }
    memcpy(out_tech_ids_host, h + o_tid, (size_t)nq * tech_limit * 8);
    memcpy(out_tech_n_host, h + o_tn, (size_t)nq * 4);
    memcpy(out_fused_ids_host, h + o_fid, (size_t)nq * max_out * 8);
    memcpy(out_fused_scores_host, h + o_fsc, (size_t)nq * max_out * 8);
    memcpy(out_fused_mask_host, h + o_fm, (size_t)nq * max_out * 4);
    memcpy(out_fused_n_host, h + o_fn, (size_t)nq * 4);
    return CDR_OK;
}
