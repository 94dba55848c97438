// tools/sanitize_driver.cc -- a plain C++ host over the C ABI (include/cadence_dense.h), no Python and no torch:
// small shapes through every kernel family, meant to run under compute-sanitizer
// (memcheck / racecheck / synccheck; profiles/r02/sanitizer.sh) where a Python + torch process would take minutes
// to start.  It is also the smallest example of a non-Python host calling the library.
//
//   g++ -O1 -std=c++17 -I include -I /usr/local/cuda/include tools/sanitize_driver.cc \
//       -L cadence_rag_b200 -lcadence_dense -L /usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/cadence_rag_b200 \
//       -o build/sanitize_driver
//   compute-sanitizer --tool memcheck build/sanitize_driver [case ...]
//
// Cases (default: all): k1 (one scan per query, k = 10/50/200, persistent over queries), k1_shared (3 queries in
// registers + deep 16-query groups), k1_gather (selective filter), bf16_scan, k2 (cluster 1/2/2-SM through
// CADENCE_K2_CLUSTER in the environment of the caller; overflow re-run through CADENCE_K2_TEST_CAP), finalize (both
// variants via batch sizes 1 and 32), filter, rrf, tech, hybrid, merge.
// Every case checks basic invariants of the results (sorted scores, ids in range) so that a silent wrong answer under
// the sanitizer is also a failure.  Exit code 0 = all cases ran and passed.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "cadence_dense.h"

#define CK(expr)                                                                                  \
    do {                                                                                          \
        int32_t rc_ = (expr);                                                                     \
        if (rc_ != CDR_OK) {                                                                      \
            fprintf(stderr, "%s:%d %s -> %d: %s\n", __FILE__, __LINE__, #expr, rc_, cdr_last_error()); \
            exit(2);                                                                              \
        }                                                                                         \
    } while (0)
#define CU(expr)                                                                                  \
    do {                                                                                          \
        cudaError_t e_ = (expr);                                                                  \
        if (e_ != cudaSuccess) {                                                                  \
            fprintf(stderr, "%s:%d %s -> %s\n", __FILE__, __LINE__, #expr, cudaGetErrorString(e_)); \
            exit(2);                                                                              \
        }                                                                                         \
    } while (0)
#define REQUIRE(cond, ...)                                                                        \
    do {                                                                                          \
        if (!(cond)) {                                                                            \
            fprintf(stderr, "%s:%d check failed: %s -- ", __FILE__, __LINE__, #cond);             \
            fprintf(stderr, __VA_ARGS__);                                                         \
            fprintf(stderr, "\n");                                                                \
            exit(3);                                                                              \
        }                                                                                         \
    } while (0)

static const int DIM = 1024;
static const uint64_t CORPUS_SEED = 20260209ull, QUERY_SEED = 20260210ull;

struct Result {
    std::vector<double> sc;
    std::vector<int64_t> id;
    std::vector<int32_t> n;
};

typedef int32_t (*dev_search_fn)(cdr_store *, const float *, int32_t, int32_t, const uint32_t *, double *, int64_t *,
                                 int32_t *, void *);

static Result run_dev(dev_search_fn fn, cdr_store *s, const float *q_dev, int nq, int k, const uint32_t *allow)
{
    double *d_sc; int64_t *d_id; int32_t *d_n;
    CU(cudaMalloc(&d_sc, (size_t)nq * k * 8)); CU(cudaMalloc(&d_id, (size_t)nq * k * 8)); CU(cudaMalloc(&d_n, (size_t)nq * 4));
    CK(fn(s, q_dev, nq, k, allow, d_sc, d_id, d_n, nullptr));
    CU(cudaDeviceSynchronize());
    Result r; r.sc.resize((size_t)nq * k); r.id.resize((size_t)nq * k); r.n.resize(nq);
    CU(cudaMemcpy(r.sc.data(), d_sc, (size_t)nq * k * 8, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(r.id.data(), d_id, (size_t)nq * k * 8, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(r.n.data(), d_n, (size_t)nq * 4, cudaMemcpyDeviceToHost));
    CU(cudaFree(d_sc)); CU(cudaFree(d_id)); CU(cudaFree(d_n));
    return r;
}

static void check_lists(const char *what, const Result &r, int nq, int k, int64_t rows, int expect_n)
{
    for (int q = 0; q < nq; ++q) {
        REQUIRE(r.n[q] == expect_n, "%s: query %d has %d results, expected %d", what, q, r.n[q], expect_n);
        for (int i = 0; i < r.n[q]; ++i) {
            const int64_t id = r.id[(size_t)q * k + i];
            REQUIRE(id >= 1 && id <= rows, "%s: id %lld out of range", what, (long long)id);
            if (i > 0) REQUIRE(r.sc[(size_t)q * k + i] <= r.sc[(size_t)q * k + i - 1], "%s: scores not descending", what);
        }
    }
}

static bool same(const Result &a, const Result &b)
{
    return a.id == b.id && a.n == b.n && memcmp(a.sc.data(), b.sc.data(), a.sc.size() * 8) == 0;
}

int main(int argc, char **argv)
{
    std::vector<std::string> want;
    for (int i = 1; i < argc; ++i) want.push_back(argv[i]);
    auto on = [&](const char *name) {
        if (want.empty()) return true;
        for (auto &w : want) if (w == name) return true;
        return false;
    };
    int32_t ndev = 0;
    CK(cdr_device_count(&ndev));
    REQUIRE(cdr_abi_version() == CDR_ABI_VERSION, "ABI version");
    const int64_t rows = 20000;            // 1250 16-row tiles: several tiles per CTA, a ragged tail for the deep kernel
    cdr_store *s = nullptr;
    CK(cdr_store_create(&s, 0, rows + 4096, DIM, CDR_STORE_FP32 | CDR_STORE_BF16));
    CK(cdr_store_append_synthetic(s, CORPUS_SEED, 0, rows, 1, 200, 1700000000000000ll, 3600000000ll, nullptr));
    CK(cdr_store_finalize(s, nullptr));
    const int NQ = 40;
    float *q_dev;
    CU(cudaMalloc(&q_dev, (size_t)NQ * DIM * 4));
    CK(cdr_synth_rows(q_dev, QUERY_SEED, 0, NQ, DIM, nullptr));
    std::vector<float> q_host((size_t)NQ * DIM);
    CU(cudaMemcpy(q_host.data(), q_dev, q_host.size() * 4, cudaMemcpyDeviceToHost));
    const int64_t words = (rows + 4096 + 31) / 32;      // bitmaps cover the capacity
    uint32_t *allow;
    CU(cudaMalloc(&allow, (size_t)words * 4));
    int64_t count = 0;

    Result base50;
    if (on("k1") || on("k1_shared") || on("k2") || on("bf16_scan")) {
        base50 = run_dev(cdr_search_exact_f32, s, q_dev, NQ, 50, nullptr);
        check_lists("k1 k=50", base50, NQ, 50, rows, 50);
    }
    if (on("k1")) {
        for (int k : {10, 200}) {
            Result r = run_dev(cdr_search_exact_f32, s, q_dev, 3, k, nullptr);
            check_lists("k1", r, 3, k, rows, k);
        }
        Result one = run_dev(cdr_search_exact_f32, s, q_dev, 1, 50, nullptr);       // cluster finalize
        REQUIRE(one.id == std::vector<int64_t>(base50.id.begin(), base50.id.begin() + 50), "single query != batch row 0");
        printf("k1 ok\n");
    }
    if (on("k1_shared")) {
        Result r = run_dev(cdr_search_exact_f32_shared, s, q_dev, NQ, 50, nullptr);   // 2 deep groups + 8 in groups of 3
        REQUIRE(same(r, base50), "shared reads differ from one scan per query");
        Result r5 = run_dev(cdr_search_exact_f32_shared, s, q_dev, 5, 50, nullptr);
        check_lists("k1_shared 5", r5, 5, 50, rows, 50);
        printf("k1_shared ok\n");
    }
    if (on("filter") || on("k1_gather") || on("k2")) {
        std::vector<uint32_t> slots(4, 0);
        slots[0] = 0x0000FFF0u;            // call slots 4..15 = rows 800..3199
        CK(cdr_filter_build(s, slots.data(), 100, 0, 0, 0, 0, 0, 0, allow, &count, nullptr));
        REQUIRE(count == 2400, "filter count %lld", (long long)count);
        printf("filter ok\n");
    }
    if (on("k1_gather")) {
        Result r = run_dev(cdr_search_exact_f32, s, q_dev, 4, 50, allow);             // gather launch serves
        check_lists("k1_gather", r, 4, 50, rows, 50);
        for (size_t i = 0; i < r.id.size(); ++i) REQUIRE(r.id[i] > 800 && r.id[i] <= 3200, "gather: id outside the filter");
        Result b = run_dev(cdr_search_scan_bf16, s, q_dev, 2, 50, allow);
        check_lists("bf16 gather", b, 2, 50, rows, 50);
        printf("k1_gather ok\n");
    }
    if (on("bf16_scan")) {
        Result r = run_dev(cdr_search_scan_bf16, s, q_dev, 6, 50, nullptr);
        check_lists("bf16_scan", r, 6, 50, rows, 50);
        int hit = 0;
        for (int q = 0; q < 6; ++q)
            for (int i = 0; i < 50; ++i)
                for (int j = 0; j < 50; ++j) hit += r.id[(size_t)q * 50 + i] == base50.id[(size_t)q * 50 + j];
        REQUIRE(hit >= 299, "bf16 scan recall %d / 300", hit);
        Result r200 = run_dev(cdr_search_scan_bf16, s, q_dev, 2, 200, nullptr);
        check_lists("bf16_scan k=200", r200, 2, 200, rows, 200);
        printf("bf16_scan ok\n");
    }
    if (on("k2")) {
        Result r = run_dev(cdr_search_batch_bf16, s, q_dev, NQ, 50, nullptr);         // 1 query tile (cluster 1)
        check_lists("k2", r, NQ, 50, rows, 50);
        const int NB = 200;                                                            // 2 query tiles: cluster of 2
        float *qb;
        CU(cudaMalloc(&qb, (size_t)NB * DIM * 4));
        CK(cdr_synth_rows(qb, QUERY_SEED, 500, NB, DIM, nullptr));
        Result rb = run_dev(cdr_search_batch_bf16, s, qb, NB, 50, nullptr);
        check_lists("k2 200", rb, NB, 50, rows, 50);
        Result rf = run_dev(cdr_search_batch_bf16, s, qb, NB, 50, allow);
        check_lists("k2 filtered", rf, NB, 50, rows, 50);
        Result rk = run_dev(cdr_search_batch_bf16, s, qb, 8, 150, nullptr);            // KC = 256
        check_lists("k2 k=150", rk, 8, 150, rows, 150);
        setenv("CADENCE_K2_TEST_CAP", "2000", 1);                                      // overflow -> device-side re-run
        Result ro = run_dev(cdr_search_batch_bf16, s, q_dev, NQ, 50, nullptr);
        unsetenv("CADENCE_K2_TEST_CAP");
        REQUIRE(same(ro, base50), "overflow re-run differs from the exact lane");
        CU(cudaFree(qb));
        printf("k2 ok\n");
    }
    if (on("rrf")) {
        const int nq = 3, L = 3;
        std::vector<int64_t> ids = {5, 7, 7, 9, 9, 5, 11,   1, 2, 3, 3, 2, 1,   42};
        std::vector<int32_t> off = {0, 2, 4, 7, 10, 13, 13, 13, 13, 14};
        std::vector<int64_t> o_ids(nq * 16); std::vector<double> o_sc(nq * 16); std::vector<uint32_t> o_m(nq * 16); std::vector<int32_t> o_n(nq);
        CK(cdr_rrf_merge_host(ids.data(), off.data(), nq, L, 60, 16, o_ids.data(), o_sc.data(), o_m.data(), o_n.data(), nullptr));
        REQUIRE(o_n[0] == 4 && o_ids[0] == 5 && o_ids[1] == 7 && o_ids[2] == 9 && o_ids[3] == 11, "rrf order");
        REQUIRE(o_sc[0] == 1.0 / 61 + 1.0 / 62, "rrf score bits");
        REQUIRE(o_n[1] == 3 && o_n[2] == 1 && o_ids[32] == 42, "rrf counts");
        printf("rrf ok\n");
    }
    cdr_tech_index *ix = nullptr;
    if (on("tech") || on("hybrid")) {
        // 3 tokens: token t owns rows r with r % 3 == t among the first 6000 rows; rank = newest call first
        std::vector<int64_t> offs = {0, 2000, 4000, 6000};
        std::vector<uint32_t> post(6000), rank(rows);
        for (int t = 0; t < 3; ++t) for (int i = 0; i < 2000; ++i) post[t * 2000 + i] = (uint32_t)(i * 3 + t);
        for (int64_t r = 0; r < rows; ++r) {
            const int64_t slot = r / 200, n_slots = rows / 200;
            rank[r] = (uint32_t)((n_slots - 1 - slot) * 200 + r % 200);
        }
        CK(cdr_tech_index_create(&ix, s, offs.data(), 3, post.data(), rank.data()));
    }
    if (on("tech")) {
        std::vector<int32_t> tok = {0, 2, -1, -1, 1, -1, -1, -1}, nt = {2, 1};
        std::vector<int64_t> o_ids(2 * 50); std::vector<int32_t> o_n(2);
        CK(cdr_tech_lane_host(ix, tok.data(), nt.data(), 2, 4, nullptr, 0, 0, 0, 0, 0, 0, 0, 50, o_ids.data(), o_n.data(), nullptr));
        REQUIRE(o_n[0] == 50 && o_n[1] == 50, "tech lane counts");
        REQUIRE(o_ids[0] > 5800 && o_ids[0] <= 6000, "tech lane head %lld is not in the newest indexed call", (long long)o_ids[0]);
        // filtered request (cluster form of the lane): call slots 0..2 = rows 0..599, newest allowed call = rows 400..599
        std::vector<uint32_t> bmap = {0x7u};
        CK(cdr_tech_lane_host(ix, tok.data(), nt.data(), 2, 4, bmap.data(), 3, 0, 0, 0, 0, 0, 0, 50, o_ids.data(), o_n.data(), nullptr));
        REQUIRE(o_n[0] == 50 && o_n[1] == 50, "filtered tech lane counts %d %d", o_n[0], o_n[1]);
        for (int i = 0; i < 100; ++i)
            REQUIRE(o_ids[i] > 400 && o_ids[i] <= 600, "filtered tech lane id %lld outside the newest allowed call", (long long)o_ids[i]);
        printf("tech ok\n");
    }
    if (on("hybrid")) {
        const int nq = 5, T = 4;
        std::vector<int32_t> tok(nq * T, -1), nt(nq, 1);
        for (int q = 0; q < nq; ++q) tok[q * T] = q % 3;
        std::vector<int64_t> bm = {10, 20, 30};
        std::vector<int32_t> bmo = {0, 3, 3, 3, 3, 3};
        const int max_out = 160;
        std::vector<int64_t> cnt(1), d_ids(nq * 50), t_ids(nq * 50), f_ids(nq * max_out);
        std::vector<double> d_sc(nq * 50), f_sc(nq * max_out);
        std::vector<int32_t> d_n(nq), t_n(nq), f_n(nq);
        std::vector<uint32_t> f_m(nq * max_out);
        for (int lane : {CDR_DENSE_LANE_EXACT_F32, CDR_DENSE_LANE_SCAN_BF16, CDR_DENSE_LANE_BATCH_BF16}) {
            cdr_filter_spec spec;
            memset(&spec, 0, sizeof(spec));
            spec.dense_lane = lane;
            CK(cdr_hybrid_retrieve_host(s, ix, &spec, q_host.data(), nq, 50, tok.data(), nt.data(), T, 50, bm.data(), bmo.data(), 60,
                                        max_out, cnt.data(), d_ids.data(), d_sc.data(), d_n.data(), t_ids.data(), t_n.data(),
                                        f_ids.data(), f_sc.data(), f_m.data(), f_n.data(), nullptr));
            REQUIRE(cnt[0] == rows, "hybrid COUNT(*) %lld", (long long)cnt[0]);
            for (int q = 0; q < nq; ++q) {
                REQUIRE(d_n[q] == 50 && t_n[q] == 50 && f_n[q] >= 50 && f_n[q] <= 103, "hybrid lane sizes");
            }
        }
        printf("hybrid ok\n");
    }
    if (on("merge")) {
        const int R = 4, nq = 3, k = 50;
        std::vector<double> sc((size_t)R * nq * k); std::vector<int64_t> id((size_t)R * nq * k); std::vector<int32_t> n(R * nq, k);
        for (int r = 0; r < R; ++r) for (int q = 0; q < nq; ++q) for (int i = 0; i < k; ++i) {
            sc[((size_t)r * nq + q) * k + i] = 1.0 - 0.001 * (i * R + r);
            id[((size_t)r * nq + q) * k + i] = 1 + i * R + r;
        }
        double *d_sc, *o_sc; int64_t *d_id, *o_id; int32_t *d_n, *o_n;
        CU(cudaMalloc(&d_sc, sc.size() * 8)); CU(cudaMalloc(&d_id, id.size() * 8)); CU(cudaMalloc(&d_n, n.size() * 4));
        CU(cudaMalloc(&o_sc, (size_t)nq * k * 8)); CU(cudaMalloc(&o_id, (size_t)nq * k * 8)); CU(cudaMalloc(&o_n, nq * 4));
        CU(cudaMemcpy(d_sc, sc.data(), sc.size() * 8, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(d_id, id.data(), id.size() * 8, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(d_n, n.data(), n.size() * 4, cudaMemcpyHostToDevice));
        CK(cdr_topk_merge(d_sc, d_id, d_n, R, nq, k, o_sc, o_id, o_n, nullptr));
        std::vector<int64_t> h_id((size_t)nq * k);
        CU(cudaMemcpy(h_id.data(), o_id, h_id.size() * 8, cudaMemcpyDeviceToHost));
        for (int i = 0; i < k; ++i) REQUIRE(h_id[i] == 1 + i, "merge order at %d: %lld", i, (long long)h_id[i]);
        // one-rank peer group: push + wait + merge on local memory
        cdr_peer_group *pg = nullptr;
        unsigned char handle[CDR_PEER_HANDLE_BYTES];
        CK(cdr_peer_group_create(&pg, 0, 0, 1, 64, 64, handle));
        CK(cdr_peer_exchange_merge(pg, d_sc, d_id, d_n, nq, k, o_sc, o_id, o_n, nullptr));
        CU(cudaDeviceSynchronize());
        CU(cudaMemcpy(h_id.data(), o_id, h_id.size() * 8, cudaMemcpyDeviceToHost));
        for (int i = 0; i < k; ++i) REQUIRE(h_id[i] == 1 + (int64_t)i * R, "peer merge (1 rank) at %d: %lld", i, (long long)h_id[i]);
        CK(cdr_peer_group_destroy(pg));
        printf("merge ok\n");
    }
    if (ix) CK(cdr_tech_index_destroy(ix));
    CU(cudaFree(allow)); CU(cudaFree(q_dev));
    CK(cdr_store_destroy(s));
    CU(cudaDeviceSynchronize());
    printf("sanitize_driver: all requested cases passed\n");
    return 0;
}
