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

namespace {

struct GroupPlan {
    bool has_filter = false, has_bm = false;
    size_t bm_words = 0;       // call-slot bitmap words
    size_t in_bm = 0;          // offset of the call-slot bitmap in the request block
    size_t x_allow = 0;        // offset of the allow bitmap in the scratch block
};

}  // namespace

// One request block for n_groups groups of consecutive queries, one filter per group.
extern "C" int32_t cdr_hybrid_retrieve_groups_host(
    cdr_store *s, cdr_tech_index *tech_index, const cdr_filter_spec *filters, const int32_t *group_offsets_host,
    int32_t n_groups, const float *q_host, int32_t nq, int32_t dense_k, const int32_t *token_ids_host,
    const int32_t *n_tokens_host, int32_t max_tokens, int32_t tech_limit, const int64_t *bm25_ids_host,
    const int32_t *bm25_offsets_host, int32_t rrf_k, int32_t max_out, int64_t *out_count_host,
    int64_t *out_dense_ids_host, double *out_dense_scores_host, int32_t *out_dense_n_host, int64_t *out_tech_ids_host,
    int32_t *out_tech_n_host, int64_t *out_fused_ids_host, double *out_fused_scores_host, uint32_t *out_fused_mask_host,
    int32_t *out_fused_n_host, void *stream)
{
    const char *fn = "cdr_hybrid_retrieve_groups_host";
    CDR_REQUIRE(s != nullptr, CDR_ERR_INVALID, "%s: store is NULL", fn);
    CDR_REQUIRE(s->finalized, CDR_ERR_STATE, "%s: store not finalized", fn);
    CDR_REQUIRE(nq >= 0 && nq <= 4096, CDR_ERR_INVALID, "%s: nq=%d outside [0,4096]", fn, nq);
    CDR_REQUIRE(n_groups >= 1 && n_groups <= 4096 && group_offsets_host != nullptr, CDR_ERR_INVALID,
                "%s: need 1 <= n_groups <= 4096 and group offsets", fn);
    CDR_REQUIRE(group_offsets_host[0] == 0 && group_offsets_host[n_groups] == nq, CDR_ERR_INVALID,
                "%s: group offsets must run from 0 to nq", fn);
    for (int gi = 0; gi < n_groups; ++gi)
        CDR_REQUIRE(group_offsets_host[gi + 1] >= group_offsets_host[gi], CDR_ERR_INVALID, "%s: group offsets not monotone", fn);
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
    if (dense && filters != nullptr) {
        for (int gi = 0; gi < n_groups; ++gi) {
            if (filters[gi].dense_lane == CDR_DENSE_LANE_EXACT_F32) continue;
            CDR_REQUIRE(filters[gi].dense_lane == CDR_DENSE_LANE_BATCH_BF16 || filters[gi].dense_lane == CDR_DENSE_LANE_SCAN_BF16,
                        CDR_ERR_INVALID, "%s: unknown dense_lane %d in group %d", fn, filters[gi].dense_lane, gi);
            CDR_REQUIRE(s->emb_bf16 != nullptr && dense_k <= 192 && s->dim % 64 == 0, CDR_ERR_UNSUPPORTED,
                        "%s: the bf16 lanes need bf16 rows, dense_k <= 192 and dim %% 64 == 0 (group %d)", fn, gi);
            CDR_REQUIRE(filters[gi].dense_lane != CDR_DENSE_LANE_SCAN_BF16 || (s->dim % 256 == 0 && s->dim <= 1024),
                        CDR_ERR_UNSUPPORTED, "%s: the bf16 scan lane needs dim in {256,512,768,1024} (group %d)", fn, gi);
        }
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

    DeviceGuard g(s->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int dim = s->dim;
    const int L = dense ? 3 : 2;
    // allow bitmaps cover the store's capacity (filter.cu): their size does not depend on the row count, which a
    // concurrent append may raise before this call takes the store lock
    const size_t words = (size_t)((s->capacity + 31) / 32);

    // ---- packed request (host -> device), packed response (device -> host), device-only scratch
    std::vector<GroupPlan> plan((size_t)n_groups);
    size_t in_q = 0, in_tok, in_nt, in_boff, in_bids, in_bm0, in_end;
    in_tok = up16(in_q + (dense ? (size_t)nq * dim * 4 : 0));
    in_nt = up16(in_tok + (tech ? (size_t)nq * max_tokens * 4 : 0));
    in_boff = up16(in_nt + (tech ? (size_t)nq * 4 : 0));
    in_bids = up16(in_boff + (bm25_offsets_host ? (size_t)(nq + 1) * 4 : 0));
    in_bm0 = up16(in_bids + (size_t)bm25_total * 8);
    size_t cur = in_bm0;
    int n_filtered = 0;
    for (int gi = 0; gi < n_groups; ++gi) {
        const cdr_filter_spec *f = filters ? &filters[gi] : nullptr;
        GroupPlan &gp = plan[(size_t)gi];
        gp.has_filter = f != nullptr && (f->call_slot_bitmap_host || f->has_date_from || f->has_date_to || f->has_tag_filter);
        gp.has_bm = gp.has_filter && f->call_slot_bitmap_host != nullptr;
        CDR_REQUIRE(!gp.has_bm || f->n_call_slots >= 0, CDR_ERR_INVALID, "%s: n_call_slots < 0 in group %d", fn, gi);
        gp.bm_words = gp.has_bm ? (size_t)((f->n_call_slots + 31) / 32) : 0;
        gp.in_bm = cur;
        cur = up16(cur + (gp.bm_words + 1) * 4);
        n_filtered += gp.has_filter ? 1 : 0;
    }
    in_end = cur;
    size_t o_cnt = in_end, o_did, o_dsc, o_dn, o_tid, o_tn, o_fid, o_fsc, o_fm, o_fn, o_end;
    o_did = up16(o_cnt + (size_t)n_groups * 8);
    o_dsc = up16(o_did + (size_t)nq * dense_k * 8);
    o_dn = up16(o_dsc + (size_t)nq * dense_k * 8);
    o_tid = up16(o_dn + (size_t)nq * 4);
    o_tn = up16(o_tid + (size_t)nq * tech_limit * 8);
    o_fid = up16(o_tn + (size_t)nq * 4);
    o_fsc = up16(o_fid + (size_t)nq * max_out * 8);
    o_fm = up16(o_fsc + (size_t)nq * max_out * 8);
    o_fn = up16(o_fm + (size_t)nq * max_out * 4);
    o_end = up16(o_fn + (size_t)nq * 4);
    cur = o_end;
    for (int gi = 0; gi < n_groups; ++gi) {
        plan[(size_t)gi].x_allow = cur;
        if (plan[(size_t)gi].has_filter) cur = up16(cur + (words + 1) * 4);
    }
    const size_t x_lids = cur;
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
    memset(h + in_bm0, 0, o_did - in_bm0);                        // call-slot bitmaps and the COUNT(*) counters start at zero
    for (int gi = 0; gi < n_groups; ++gi) {
        const GroupPlan &gp = plan[(size_t)gi];
        if (gp.has_bm && gp.bm_words) memcpy(h + gp.in_bm, filters[gi].call_slot_bitmap_host, gp.bm_words * 4);
    }
    CDR_CUDA(cudaMemcpyAsync(d, h, o_did, cudaMemcpyHostToDevice, st));

    int64_t *d_tid = (int64_t *)(d + o_tid);
    int32_t *d_tn = (int32_t *)(d + o_tn);
    if (!tech) {
        CDR_CUDA(cudaMemsetAsync(d_tid, 0xFF, (size_t)nq * tech_limit * 8, st));
        CDR_CUDA(cudaMemsetAsync(d_tn, 0, (size_t)nq * 4, st));
    }
    int rc;
    for (int gi = 0; gi < n_groups; ++gi) {
        const int q0 = group_offsets_host[gi], gq = group_offsets_host[gi + 1] - q0;
        if (gq == 0) continue;
        const GroupPlan &gp = plan[(size_t)gi];
        const cdr_filter_spec *f = gp.has_filter ? &filters[gi] : nullptr;
        const uint32_t *d_bm = gp.has_bm ? (const uint32_t *)(d + gp.in_bm) : nullptr;
        const int64_t n_slots = gp.has_bm ? f->n_call_slots : 0;
        const int f_from = f ? f->has_date_from : 0, f_to = f ? f->has_date_to : 0, f_tags = f ? f->has_tag_filter : 0;
        const int64_t t_from = f ? f->date_from_us : 0, t_to = f ? f->date_to_us : 0;
        const uint64_t tag_any = f ? f->tag_any : 0;
        // tech_tokens lane first: short, and independent of the dense lane
        if (tech) {
            rc = cdr_tech_lane_launch(tech_index, (const int32_t *)(d + in_tok) + (size_t)q0 * max_tokens,
                                      (const int32_t *)(d + in_nt) + q0, gq, max_tokens, d_bm, n_slots, f_from, t_from, f_to,
                                      t_to, f_tags, tag_any, tech_limit, d_tid + (size_t)q0 * tech_limit, d_tn + q0, st);
            if (rc != CDR_OK) return rc;
        }
        // dense lane: WHERE <filters> AND embedding IS NOT NULL, ORDER BY embedding <=> q LIMIT k
        const uint32_t *allow = s->any_invalid ? s->valid : nullptr;
        if (gp.has_filter) {
            uint32_t *d_allow = (uint32_t *)(d + gp.x_allow);
            rc = cdr_filter_launch(s, d_bm, n_slots, f_from, t_from, f_to, t_to, f_tags, tag_any, d_allow,
                                   (unsigned long long *)(d + o_cnt) + gi, st);
            if (rc != CDR_OK) return rc;
            allow = d_allow;
        }
        if (dense) {
            // dense_lane == CDR_DENSE_LANE_BATCH_BF16: the group's planner mode is "ann" (the reference would walk its
            // HNSW index, app/retrieve.py:291-298) and the caller prefers the tensor-core lane for this batch
            const bool ann = filters != nullptr && filters[gi].dense_lane == CDR_DENSE_LANE_BATCH_BF16;
            const float *gq_dev = (const float *)(d + in_q) + (size_t)q0 * dim;
            double *g_sc = (double *)(d + o_dsc) + (size_t)q0 * dense_k;
            int64_t *g_id = (int64_t *)(d + o_did) + (size_t)q0 * dense_k;
            int32_t *g_n = (int32_t *)(d + o_dn) + q0;
            const bool ann_scan = filters != nullptr && filters[gi].dense_lane == CDR_DENSE_LANE_SCAN_BF16;
            if (ann) rc = cdr_batch_bf16_launch(s, ws, gq_dev, gq, dense_k, allow, g_sc, g_id, g_n, st);
            else if (ann_scan) rc = cdr_bf16_scan_launch(s, ws, gq_dev, gq, allow, dense_k, g_sc, g_id, g_n, st);
            else rc = cdr_exact_scan_launch(s, ws, gq_dev, gq, allow, dense_k, g_sc, g_id, g_n, st, /*share_reads=*/true);
            if (rc != CDR_OK) return rc;
        }
    }

    // lanes -> RRF, all queries at once
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
        for (int gi = 0; gi < n_groups; ++gi) {
            unsigned long long c;
            memcpy(&c, h + o_cnt + (size_t)gi * 8, 8);
            out_count_host[gi] = plan[(size_t)gi].has_filter ? (int64_t)c : n_valid_now;
        }
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

// All queries share one filter: a single group.
extern "C" int32_t cdr_hybrid_retrieve_host(
    cdr_store *s, cdr_tech_index *tech_index, const cdr_filter_spec *filter, const float *q_host, int32_t nq,
    int32_t dense_k, const int32_t *token_ids_host, const int32_t *n_tokens_host, int32_t max_tokens,
    int32_t tech_limit, const int64_t *bm25_ids_host, const int32_t *bm25_offsets_host, int32_t rrf_k,
    int32_t max_out, int64_t *out_count_host, int64_t *out_dense_ids_host, double *out_dense_scores_host,
    int32_t *out_dense_n_host, int64_t *out_tech_ids_host, int32_t *out_tech_n_host, int64_t *out_fused_ids_host,
    double *out_fused_scores_host, uint32_t *out_fused_mask_host, int32_t *out_fused_n_host, void *stream)
{
    const int32_t offsets[2] = {0, nq};
    return cdr_hybrid_retrieve_groups_host(s, tech_index, filter, offsets, 1, q_host, nq, dense_k, token_ids_host,
                                           n_tokens_host, max_tokens, tech_limit, bm25_ids_host, bm25_offsets_host, rrf_k,
                                           max_out, out_count_host, out_dense_ids_host, out_dense_scores_host,
                                           out_dense_n_host, out_tech_ids_host, out_tech_n_host, out_fused_ids_host,
                                           out_fused_scores_host, out_fused_mask_host, out_fused_n_host, stream);
}
