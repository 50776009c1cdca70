/* oracle_fast.c — TEST / BENCH INFRASTRUCTURE ONLY (never linked into the product).
 *
 * orc_match_l2_fast: the same exact brute-force L2 1-NN as orc_match_l2 (oracle.c), written the way a tuned CPU
 * implementation would be, so that bench.py's CPU arm is not a strawman: |a-b|^2 = |a|^2 + |b|^2 - 2 a.b with the dot
 * product on 16-bit lanes (AVX2 vpmaddwd: 16 exact u8*u8 MACs per instruction), four query rows per pass over the
 * train set.  Results are bit-identical to orc_match_l2 (integer arithmetic, lowest train index on ties); machines
 * without AVX2 fall back to orc_match_l2.  tests/test_cpu_oracle.py checks fast == plain.
 * Stands in for FlannBasedMatcher().match (M/MosaicWithoutPos.cpp:5108-5110), which is approximate. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>
#include <immintrin.h>
#include "oracle.h"

__attribute__((target("avx2")))
static inline int32_t hsum256(__m256i v)
{
    __m128i s = _mm_add_epi32(_mm256_castsi256_si128(v), _mm256_extracti128_si256(v, 1));
    s = _mm_add_epi32(s, _mm_shuffle_epi32(s, 0x4e));
    s = _mm_add_epi32(s, _mm_shuffle_epi32(s, 0xb1));
    return _mm_cvtsi128_si32(s);
}

__attribute__((target("avx2")))
static void match_avx2(const int16_t* A16, const int32_t* na2, int na, const int16_t* B16, const int32_t* nb2, int nb,
                       int32_t* train_idx, int32_t* d2)
{
    /* dim == 128: a row is 8 vectors of 16 int16 */
#pragma omp parallel for schedule(static)
    for (int i0 = 0; i0 < na; i0 += 4) {
        const int ni = na - i0 < 4 ? na - i0 : 4;
        __m256i a[4][8];
        for (int r = 0; r < 4; r++)
            for (int k = 0; k < 8; k++)
                a[r][k] = _mm256_loadu_si256((const __m256i*)(A16 + (size_t)(i0 + (r < ni ? r : 0)) * 128 + 16 * k));
        int32_t best[4] = {INT_MAX, INT_MAX, INT_MAX, INT_MAX}, bj[4] = {-1, -1, -1, -1};
        for (int j = 0; j < nb; j++) {
            const __m256i* b = (const __m256i*)(B16 + (size_t)j * 128);
            __m256i acc0 = _mm256_setzero_si256(), acc1 = acc0, acc2 = acc0, acc3 = acc0;
            for (int k = 0; k < 8; k++) {
                const __m256i bv = _mm256_loadu_si256(b + k);
                acc0 = _mm256_add_epi32(acc0, _mm256_madd_epi16(a[0][k], bv));
                acc1 = _mm256_add_epi32(acc1, _mm256_madd_epi16(a[1][k], bv));
                acc2 = _mm256_add_epi32(acc2, _mm256_madd_epi16(a[2][k], bv));
                acc3 = _mm256_add_epi32(acc3, _mm256_madd_epi16(a[3][k], bv));
            }
            const int32_t dot[4] = {hsum256(acc0), hsum256(acc1), hsum256(acc2), hsum256(acc3)};
            for (int r = 0; r < 4; r++) {
                const int32_t d = na2[i0 + (r < ni ? r : 0)] + nb2[j] - 2 * dot[r];
                if (d < best[r]) { best[r] = d; bj[r] = j; }
            }
        }
        for (int r = 0; r < ni; r++) { train_idx[i0 + r] = bj[r]; d2[i0 + r] = best[r]; }
    }
}

void orc_match_l2_fast(const uint8_t* A, int na, const uint8_t* B, int nb, int dim, int32_t* train_idx, int32_t* d2)
{
    if (dim != 128 || na <= 0 || nb <= 0 || !__builtin_cpu_supports("avx2")) { orc_match_l2(A, na, B, nb, dim, train_idx, d2); return; }
    int16_t* A16 = (int16_t*)malloc((size_t)na * 128 * 2); int16_t* B16 = (int16_t*)malloc((size_t)nb * 128 * 2);
    int32_t* na2 = (int32_t*)malloc((size_t)na * 4); int32_t* nb2 = (int32_t*)malloc((size_t)nb * 4);
    for (int i = 0; i < na; i++) { int32_t s = 0; for (int k = 0; k < 128; k++) { const int v = A[(size_t)i * 128 + k]; A16[(size_t)i * 128 + k] = (int16_t)v; s += v * v; } na2[i] = s; }
    for (int i = 0; i < nb; i++) { int32_t s = 0; for (int k = 0; k < 128; k++) { const int v = B[(size_t)i * 128 + k]; B16[(size_t)i * 128 + k] = (int16_t)v; s += v * v; } nb2[i] = s; }
    match_avx2(A16, na2, na, B16, nb2, nb, train_idx, d2);
    free(A16); free(B16); free(na2); free(nb2);
}
