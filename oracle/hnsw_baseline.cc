// oracle/hnsw_baseline.cc -- TEST/BENCH INFRASTRUCTURE ONLY (CPU). Never used by the product.
//
// CPU baseline for the reference's mode="ann" (app/retrieve.py:290-298 turns on pgvector's HNSW
// index with hnsw.ef_search = 80; DDL: USING hnsw (embedding vector_cosine_ops) WITH
// (m = 16, ef_construction = 64), alembic/versions/0001_initial_schema.py:98-102).
// pgvector is not vendored in the reference and cannot run in this image, so this is a
// restatement of the published HNSW algorithm (Malkov & Yashunin 2018, Alg. 1-5) with pgvector's
// parameters: level multiplier 1/ln(m), 2*m links on layer 0, vector_cosine_ops = vectors
// normalised once and compared by negative inner product, neighbour selection by the simple
// closest-first heuristic with pruning.  It is labelled "restated, not Postgres" wherever its
// numbers are reported; it exists to time a CPU graph walk next to the GPU brute-force lane and
// to report its recall against the exact oracle.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <omp.h>
#include <queue>
#include <random>
#include <vector>

namespace {

struct Hnsw {
    int dim = 0, m = 16, efc = 64;
    int64_t n = 0;
    const float *x = nullptr;          // [n, dim], L2-normalised
    std::vector<int> level;            // per node
    std::vector<std::vector<std::vector<int>>> links;   // [node][layer] -> neighbours
    int entry = -1, max_level = -1;
    // visited marks live in a per-thread scratch so queries can run concurrently (the build is serial)
    struct Scratch {
        std::vector<uint32_t> visited;
        uint32_t epoch = 0;
    };
    Scratch build_scratch;

    float dist(const float *a, const float *b) const {
        float s0 = 0, s1 = 0, s2 = 0, s3 = 0;
        for (int i = 0; i < dim; i += 4) {
            s0 += a[i] * b[i]; s1 += a[i + 1] * b[i + 1]; s2 += a[i + 2] * b[i + 2]; s3 += a[i + 3] * b[i + 3];
        }
        return -(s0 + s1 + s2 + s3);   // negative inner product (smaller = closer)
    }

    typedef std::pair<float, int> Cand;

    // Alg. 2: best-first search on one layer with beam `ef`; returns up to ef closest (max-heap order)
    std::vector<Cand> search_layer(Scratch &sc, const float *q, std::vector<Cand> entry_pts, int ef, int layer) {
        std::vector<uint32_t> &visited = sc.visited;
        if (visited.size() != (size_t)n) { visited.assign(n, 0); sc.epoch = 0; }
        if (++sc.epoch == 0) { std::fill(visited.begin(), visited.end(), 0); sc.epoch = 1; }
        const uint32_t epoch = sc.epoch;
        std::priority_queue<Cand, std::vector<Cand>, std::greater<Cand>> cand;   // closest first
        std::priority_queue<Cand> best;                                           // farthest on top
        for (auto &e : entry_pts) { visited[e.second] = epoch; cand.push(e); best.push(e); }
        while (!cand.empty()) {
            Cand c = cand.top(); cand.pop();
            if (c.first > best.top().first && (int)best.size() >= ef) break;
            for (int nb : links[c.second][layer]) {
                if (visited[nb] == epoch) continue;
                visited[nb] = epoch;
                float d = dist(q, x + (int64_t)nb * dim);
                if ((int)best.size() < ef || d < best.top().first) {
                    cand.push({d, nb}); best.push({d, nb});
                    if ((int)best.size() > ef) best.pop();
                }
            }
        }
        std::vector<Cand> out;
        while (!best.empty()) { out.push_back(best.top()); best.pop(); }
        std::reverse(out.begin(), out.end());   // closest first
        return out;
    }

    // Alg. 4 (heuristic): keep a candidate only if it is closer to the base than to every kept one
    std::vector<int> select_neighbours(const std::vector<Cand> &cands, int mmax) {
        std::vector<int> kept;
        for (auto &c : cands) {
            if ((int)kept.size() >= mmax) break;
            bool ok = true;
            for (int k : kept)
                if (dist(x + (int64_t)c.second * dim, x + (int64_t)k * dim) < c.first) { ok = false; break; }
            if (ok) kept.push_back(c.second);
        }
        return kept;
    }

    void build(const float *rows, int64_t n_, int dim_, int m_, int efc_, uint64_t seed) {
        x = rows; n = n_; dim = dim_; m = m_; efc = efc_;
        level.resize(n); links.resize(n);
        Scratch &sc = build_scratch;
        std::mt19937_64 rng(seed);
        std::uniform_real_distribution<double> uni(0.0, 1.0);
        const double ml = 1.0 / std::log((double)m);
        for (int64_t i = 0; i < n; ++i) {
            int lv = (int)(-std::log(std::max(uni(rng), 1e-300)) * ml);
            level[i] = lv;
            links[i].resize(lv + 1);
            const float *q = x + i * dim;
            if (entry < 0) { entry = (int)i; max_level = lv; continue; }
            std::vector<Cand> ep = {{dist(q, x + (int64_t)entry * dim), entry}};
            for (int l = max_level; l > lv; --l) ep = {search_layer(sc, q, ep, 1, l)[0]};
            for (int l = std::min(lv, max_level); l >= 0; --l) {
                std::vector<Cand> w = search_layer(sc, q, ep, efc, l);
                const int mmax = l == 0 ? 2 * m : m;
                std::vector<int> nb = select_neighbours(w, m);
                links[i][l] = nb;
                // back-links: the neighbours' lists are distinct, so their re-pruning runs in parallel
                // (the result does not depend on the thread count)
#pragma omp parallel for schedule(dynamic, 1)
                for (int t0 = 0; t0 < (int)nb.size(); ++t0) {
                    const int o = nb[t0];
                    auto &ol = links[o][l];
                    ol.push_back((int)i);
                    if ((int)ol.size() > mmax) {
                        std::vector<Cand> oc;
                        for (int t : ol) oc.push_back({dist(x + (int64_t)o * dim, x + (int64_t)t * dim), t});
                        std::sort(oc.begin(), oc.end());
                        ol = select_neighbours(oc, mmax);
                    }
                }
                ep = w;
            }
            if (lv > max_level) { max_level = lv; entry = (int)i; }
        }
    }

    int search(Scratch &sc, const float *q, int k, int ef, int64_t *out_rows, float *out_sim) {
        if (entry < 0) return 0;
        std::vector<Cand> ep = {{dist(q, x + (int64_t)entry * dim), entry}};
        for (int l = max_level; l > 0; --l) ep = {search_layer(sc, q, ep, 1, l)[0]};
        std::vector<Cand> w = search_layer(sc, q, ep, std::max(ef, k), 0);
        int m_out = std::min<int>(k, (int)w.size());
        for (int i = 0; i < m_out; ++i) { out_rows[i] = w[i].second; out_sim[i] = -w[i].first; }
        return m_out;
    }
};

}  // namespace

extern "C" {
void *orc_hnsw_build(const float *rows_normalised, int64_t n, int dim, int m, int ef_construction, uint64_t seed)
{
    Hnsw *h = new Hnsw();
    h->build(rows_normalised, n, dim, m, ef_construction, seed);
    return h;
}
// q must be L2-normalised by the caller; returns #results; rows are 0-based row indices
int orc_hnsw_search(void *h, const float *q, int k, int ef_search, int64_t *out_rows, float *out_sim)
{
    Hnsw *g = static_cast<Hnsw *>(h);
    return g->search(g->build_scratch, q, k, ef_search, out_rows, out_sim);
}
// nq independent queries ([nq, dim], L2-normalised) on `nthreads` OpenMP threads (<= 0: all cores), one
// scratch per thread -- the analogue of nq concurrent Postgres backends walking the same index.
// out_rows / out_sim are [nq, k] (unused slots -1 / 0); out_n [nq].
void orc_hnsw_search_batch(void *h, const float *qs, int nq, int k, int ef_search, int64_t *out_rows, float *out_sim,
                           int *out_n, int nthreads)
{
    Hnsw *g = static_cast<Hnsw *>(h);
    if (nthreads <= 0) nthreads = omp_get_num_procs();
#pragma omp parallel num_threads(nthreads)
    {
        Hnsw::Scratch sc;
#pragma omp for schedule(dynamic, 4)
        for (int i = 0; i < nq; ++i) {
            for (int j = 0; j < k; ++j) { out_rows[(int64_t)i * k + j] = -1; out_sim[(int64_t)i * k + j] = 0.f; }
            out_n[i] = g->search(sc, qs + (int64_t)i * g->dim, k, ef_search, out_rows + (int64_t)i * k, out_sim + (int64_t)i * k);
        }
    }
}
void orc_hnsw_free(void *h) { delete static_cast<Hnsw *>(h); }
}
