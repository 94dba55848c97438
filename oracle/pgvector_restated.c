/*
 * oracle/pgvector_restated.c -- TEST INFRASTRUCTURE ONLY (CPU). Never imported by the product
 * path; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it.
 *
 * PARITY UNPINNED: the arithmetic of the reference's dense lane lives in a third-party
 * dependency that is NOT vendored under /root/reference: pgvector 0.8.1 (pinned at
 * app/config.py:8, docker-compose.yml:5, enforced at app/db.py:56-61).  No reference test
 * or fixture pins dense results (tests/conftest.py:85,95 and
 * eval/run_real_regression_gate.py:141 disable the dense lane).  This file restates the
 * PUBLISHED pgvector 0.8.1 algorithm (src/vector.c: VectorCosineSimilarity /
 * cosine_distance) and the SQL the reference wraps around it:
 *
 *   app/retrieve.py:339-351   SELECT ..., (1 - (embedding <=> :q)) AS score
 *                             FROM chunks WHERE <filters> AND embedding IS NOT NULL
 *                             ORDER BY embedding <=> :q LIMIT :limit
 *   app/retrieve.py:374-386   same over artifact_chunks
 *
 * pgvector's kernel (restated, not copied): three fp32 accumulators (a.b, a.a, b.b) over the
 * dim elements, then  sim = (double)ab / sqrt((double)aa * (double)bb) , clamp to [-1,1],
 * distance = 1.0 - sim (float8).  pgvector builds that loop with
 *   -ftree-vectorize -fassociative-math -fno-signed-zeros -fno-trapping-math
 * and target_clones("default","fma"), so the fp32 summation order is compiler-defined; build
 * THIS file with the same flags (oracle/Makefile does).  The exact plan is a sequential scan
 * feeding a bounded top-N sort on the float8 distance; NaN distances (zero-norm vectors) sort
 * last (PostgreSQL float8 ordering).  The reference SQL has no tiebreak key; this engine's
 * contract (BASELINE.json north_star) breaks ties by chunk_id ascending, so the oracle does too.
 *
 * Two scoring variants are exported:
 *   pgv32 : the fp32-accumulate loop above           ("what pgvector would say")
 *   f64   : the same formula with fp64 accumulators  (ground-truth order for the id-list test)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#if defined(__x86_64__) && defined(__gnu_linux__) && defined(__has_attribute)
#if __has_attribute(target_clones)
#define ORC_TARGET_CLONES __attribute__((target_clones("default", "fma")))
#endif
#endif
#ifndef ORC_TARGET_CLONES
#define ORC_TARGET_CLONES
#endif

/* pgvector 0.8.1 src/vector.c VectorCosineSimilarity (restated). */
ORC_TARGET_CLONES static double cosine_similarity_pgv32(int dim, const float *ax, const float *bx)
{
    float ab = 0.0f, aa = 0.0f, bb = 0.0f;
    for (int i = 0; i < dim; i++) {
        ab += ax[i] * bx[i];
        aa += ax[i] * ax[i];
        bb += bx[i] * bx[i];
    }
    return (double)ab / sqrt((double)aa * (double)bb);
}

static double cosine_similarity_f64(int dim, const float *ax, const float *bx)
{
    double ab = 0.0, aa = 0.0, bb = 0.0;
    for (int i = 0; i < dim; i++) {
        double a = ax[i], b = bx[i];
        ab += a * b;
        aa += a * a;
        bb += b * b;
    }
    return ab / sqrt(aa * bb);
}

/* pgvector cosine_distance: clamp, 1 - sim.  NaN propagates (0/0 for a zero vector). */
static inline double distance_from_similarity(double sim)
{
    if (sim > 1.0) sim = 1.0;
    else if (sim < -1.0) sim = -1.0;
    return 1.0 - sim;
}

double orc_cosine_distance_pgv32(const float *a, const float *b, int dim)
{
    return distance_from_similarity(cosine_similarity_pgv32(dim, a, b));
}

double orc_cosine_distance_f64(const float *a, const float *b, int dim)
{
    return distance_from_similarity(cosine_similarity_f64(dim, a, b));
}

/* ORDER BY distance ASC (NaN last), then id ASC. Returns <0 when x sorts before y. */
typedef struct { double dist; int64_t id; } cand_t;

static inline int cand_before(const cand_t *x, const cand_t *y)
{
    int xn = isnan(x->dist), yn = isnan(y->dist);
    if (xn != yn) return yn;            /* non-NaN first */
    if (!xn && x->dist != y->dist) return x->dist < y->dist;
    return x->id < y->id;
}

static int cand_cmp(const void *pa, const void *pb)
{
    const cand_t *x = (const cand_t *)pa, *y = (const cand_t *)pb;
    if (cand_before(x, y)) return -1;
    if (cand_before(y, x)) return 1;
    return 0;
}

/* Bounded insertion buffer: keeps the `k` best candidates seen, sorted. */
typedef struct { cand_t *v; int n; int k; } topn_t;

static inline void topn_push(topn_t *t, cand_t c)
{
    if (t->n == t->k && !cand_before(&c, &t->v[t->n - 1])) return;
    int pos = (t->n < t->k) ? t->n++ : t->n - 1;
    while (pos > 0 && cand_before(&c, &t->v[pos - 1])) { t->v[pos] = t->v[pos - 1]; pos--; }
    t->v[pos] = c;
}

/*
 * Exact scan: the SQL at app/retrieve.py:339-351 over an in-memory table.
 *   q[dim], x[n*dim] fp32, ids[n] (NULL => id = row+1, BIGSERIAL),
 *   allow: optional bitmap, bit r set => row r passes WHERE <filters> AND embedding IS NOT NULL,
 *   variant 0 = pgv32, 1 = f64;  nthreads <= 0 => all cores.
 * Writes out_score[k] = 1 - distance (app/retrieve.py:343), out_id[k]; returns #rows written
 * (SQL LIMIT semantics: short when fewer rows survive).
 */
int orc_exact_scan(const float *q, const float *x, const int64_t *ids, int64_t n, int dim,
                   const uint32_t *allow, int k, int variant, int nthreads,
                   double *out_score, int64_t *out_id)
{
    if (k <= 0 || n <= 0) return 0;
    int nt = 1;
#ifdef _OPENMP
    nt = nthreads > 0 ? nthreads : omp_get_max_threads();
#endif
    if (nt > n) nt = (int)n;
    cand_t *all = (cand_t *)malloc(sizeof(cand_t) * (size_t)nt * (size_t)k);
    int *cnt = (int *)calloc((size_t)nt, sizeof(int));
#pragma omp parallel num_threads(nt)
    {
        int tid = 0;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#endif
        topn_t t = { all + (size_t)tid * k, 0, k };
        int64_t lo = n * tid / nt, hi = n * (tid + 1) / nt;
        for (int64_t r = lo; r < hi; ++r) {
            if (allow && !((allow[r >> 5] >> (r & 31)) & 1u)) continue;
            const float *row = x + r * (int64_t)dim;
            double sim = variant ? cosine_similarity_f64(dim, row, q)
                                 : cosine_similarity_pgv32(dim, row, q);
            cand_t c = { distance_from_similarity(sim), ids ? ids[r] : r + 1 };
            topn_push(&t, c);
        }
        cnt[tid] = t.n;
    }
    int total = 0;
    for (int t = 0; t < nt; ++t) {
        if (t * k != total) memmove(all + total, all + (size_t)t * k, sizeof(cand_t) * cnt[t]);
        total += cnt[t];
    }
    qsort(all, (size_t)total, sizeof(cand_t), cand_cmp);
    int m = total < k ? total : k;
    for (int i = 0; i < m; ++i) {
        out_score[i] = 1.0 - all[i].dist;   /* SELECT (1 - (embedding <=> q)) AS score */
        out_id[i] = all[i].id;
    }
    free(all); free(cnt);
    return m;
}

/* All scores (1 - distance) for every row; used by the near-tie analysis in tests. */
void orc_all_scores(const float *q, const float *x, int64_t n, int dim, int variant,
                    double *out_score)
{
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < n; ++r) {
        const float *row = x + r * (int64_t)dim;
        double sim = variant ? cosine_similarity_f64(dim, row, q)
                             : cosine_similarity_pgv32(dim, row, q);
        out_score[r] = 1.0 - distance_from_similarity(sim);
    }
}

/*
 * bf16-valued corpus truth: rows are bf16 bit patterns widened to fp32,
 * query fp32, fp64 accumulate.  Same ordering rules.
 */
int orc_exact_scan_bf16rows(const float *q, const uint16_t *xb, const int64_t *ids, int64_t n,
                            int dim, const uint32_t *allow, int k, double *out_score,
                            int64_t *out_id)
{
    if (k <= 0 || n <= 0) return 0;
    cand_t *all = (cand_t *)malloc(sizeof(cand_t) * (size_t)n);
    int64_t m = 0;
    double *sc = (double *)malloc(sizeof(double) * (size_t)n);
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < n; ++r) {
        const uint16_t *row = xb + r * (int64_t)dim;
        double ab = 0.0, aa = 0.0, bb = 0.0;
        for (int i = 0; i < dim; ++i) {
            uint32_t u = (uint32_t)row[i] << 16;
            float f; memcpy(&f, &u, 4);
            double a = f, b = q[i];
            ab += a * b; aa += a * a; bb += b * b;
        }
        sc[r] = distance_from_similarity(ab / sqrt(aa * bb));
    }
    for (int64_t r = 0; r < n; ++r) {
        if (allow && !((allow[r >> 5] >> (r & 31)) & 1u)) continue;
        all[m].dist = sc[r]; all[m].id = ids ? ids[r] : r + 1; m++;
    }
    qsort(all, (size_t)m, sizeof(cand_t), cand_cmp);
    int w = m < k ? (int)m : k;
    for (int i = 0; i < w; ++i) { out_score[i] = 1.0 - all[i].dist; out_id[i] = all[i].id; }
    free(all); free(sc);
    return w;
}

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
