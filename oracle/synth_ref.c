/*
 * oracle/synth_ref.c -- TEST INFRASTRUCTURE ONLY (CPU). Never imported by the product path.
 *
 * CPU statement of the synthetic corpus/query generator used by the parity tests and by
 * bench.py (SURVEY.md section 8(d) "Synthetic inputs", BASELINE.md "Data").  The CUDA
 * generator in cadence_rag_b200/csrc/store.cu must produce bit-identical fp32 rows; this file
 * is the independent restatement it is checked against (tests/test_gpu_store.py).
 *
 * Specification (chosen so that CPU and GPU agree bit for bit -- no transcendental
 * functions, only integer arithmetic plus one IEEE sqrt, one IEEE divide and one IEEE
 * multiply per element):
 *
 *   w[0..3]  = Philox-4x32-10(key = (seed_lo, seed_hi),
 *                             ctr = (row_lo, row_hi, j/4, 0))        for elements j..j+3
 *   s_j      = byte0(w) + byte1(w) + byte2(w) + byte3(w) - 510       (Irwin-Hall(4) of bytes:
 *                                                                     zero-mean, ~Gaussian)
 *   S        = sum_j s_j^2                                           (exact, 64-bit integer)
 *   inv      = 1.0f / sqrtf((float)S)                                (inv = 0 when S == 0)
 *   x[row,j] = (float)s_j * inv                                      (L2-normalised in fp32, like
 *                                                                     the embedding gateway's
 *                                                                     _normalize, reference
 *                                                                     P620_..._RUNBOOK.md:510-513)
 *   bf16     = round-to-nearest-even of x[row,j]
 *
 * `row` is the GLOBAL row index, so any sharding of the corpus sees identical data.
 * Compile with plain -O2 (no -ffast-math / -fassociative-math).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define PHILOX_M0 0xD2511F53u
#define PHILOX_M1 0xCD9E8D57u
#define PHILOX_W0 0x9E3779B9u
#define PHILOX_W1 0xBB67AE85u

static inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                 uint32_t k0, uint32_t k1, uint32_t out[4])
{
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)PHILOX_M0 * c0;
        uint64_t p1 = (uint64_t)PHILOX_M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += PHILOX_W0; k1 += PHILOX_W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static inline int32_t ih4(uint32_t w)
{
    return (int32_t)((w & 0xFFu) + ((w >> 8) & 0xFFu) + ((w >> 16) & 0xFFu) + (w >> 24)) - 510;
}

/* Known-answer hook for the Philox core (Random123 KAT vectors are checked in tests). */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1], out);
}

/* One row. dim must be a multiple of 4. */
static void synth_row(float *dst, uint64_t seed, int64_t row, int dim)
{
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    uint32_t r0 = (uint32_t)(uint64_t)row, r1 = (uint32_t)((uint64_t)row >> 32);
    int64_t sumsq = 0;
    int32_t *tmp = (int32_t *)dst; /* same size as float; converted in place below */
    for (int j = 0; j < dim; j += 4) {
        uint32_t w[4];
        philox4x32_10(r0, r1, (uint32_t)(j >> 2), 0u, k0, k1, w);
        for (int t = 0; t < 4; ++t) {
            int32_t s = ih4(w[t]);
            tmp[j + t] = s;
            sumsq += (int64_t)s * s;
        }
    }
    float inv = 0.0f;
    if (sumsq > 0) inv = 1.0f / sqrtf((float)sumsq);
    for (int j = 0; j < dim; ++j) {
        int32_t s = tmp[j];
        dst[j] = (float)s * inv;
    }
}

void orc_synth_rows(float *dst, uint64_t seed, int64_t first_row, int64_t n, int dim)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i)
        synth_row(dst + i * (int64_t)dim, seed, first_row + i, dim);
}

/* fp32 -> bf16 bits, round-to-nearest-even, NaN quieted (matches __float2bfloat16_rn). */
void orc_f32_to_bf16(const float *src, uint16_t *dst, int64_t n)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        uint32_t u;
        memcpy(&u, &src[i], 4);
        if ((u & 0x7FFFFFFFu) > 0x7F800000u) { dst[i] = 0x7FFF; continue; }
        uint32_t lsb = (u >> 16) & 1u;
        u += 0x7FFFu + lsb;
        dst[i] = (uint16_t)(u >> 16);
    }
}

void orc_bf16_to_f32(const uint16_t *src, float *dst, int64_t n)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        uint32_t u = (uint32_t)src[i] << 16;
        memcpy(&dst[i], &u, 4);
    }
}

/* Per-call tag mask of the synthetic corpus: two Philox-chosen tags out of 16
 * ("tag_bits = 2 random of 16 tags per call"). ctr = (slot_lo, slot_hi, 0, 1). */
uint64_t orc_synth_tag_bits(uint64_t seed, int64_t call_slot)
{
    uint32_t w[4];
    philox4x32_10((uint32_t)(uint64_t)call_slot, (uint32_t)((uint64_t)call_slot >> 32), 0u, 1u,
                  (uint32_t)seed, (uint32_t)(seed >> 32), w);
    return (1ull << (w[0] & 15u)) | (1ull << (w[1] & 15u));
}
