// ransac_math.cuh — per-hypothesis arithmetic of K4 (RANSAC homography), written so that every float
// operation happens in the order of the reference (M/matrix.h, M/LeastSquare.h, M/mosaicimage.h; see
// SURVEY.md Appendix C).  Compiled for the device with -fmad=false (no FMA contraction) and IEEE
// division/sqrt, and — for CPU unit tests only (tests/host_harness) — for the host with
// -ffp-contract=off, so the same source can be checked against the oracle without a GPU.
//
// Two paths:
//   fast  : fully unrolled, register resident, exploits the fixed sparsity of the 4-point systems and an
//           in-place Gauss-Jordan that is bit-identical to the reference's augmented-matrix version as
//           long as (a) every pivot is the diagonal entry, (b) every skipped multiplier is exactly zero
//           and (c) all results are finite.  It reports need_slow when an assumption does not hold.
//   slow  : generic restatement with the reference's pivot search / row re-ordering quirks
//           (M/matrix.h:147-296), used for the rare tuples the fast path rejects.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define UAVM_HD __host__ __device__ __forceinline__
#define UAVM_HD_NOINLINE __host__ __device__ __noinline__
#else
#define UAVM_HD inline
#define UAVM_HD_NOINLINE inline
#endif

namespace uavm {
namespace rmath {

// MSVC rand() LCG, the sample stream contract (replaces srand(time(0)), M/mosaicimage.h:1777).
constexpr uint32_t kLcgA = 214013u, kLcgC = 2531011u;
UAVM_HD uint32_t lcg_next(uint32_t& s) { s = s * kLcgA + kLcgC; return (s >> 16) & 0x7fffu; }
// state after k steps from s: affine map composition by binary powering
UAVM_HD uint32_t lcg_jump(uint32_t s, uint32_t k) {
    uint32_t a = kLcgA, c = kLcgC;          // current power-of-two step map x -> a x + c
    uint32_t ra = 1u, rc = 0u;              // accumulated map
    while (k) {
        if (k & 1u) { ra = ra * a; rc = rc * a + c; }
        c = c * a + c; a = a * a;
        k >>= 1;
    }
    return ra * s + rc;
}
// draws group g (4 consecutive rand() % n); returns true if the four indices are distinct
UAVM_HD bool draw_group(uint32_t seed, uint32_t g, int n, int idx[4]) {
    uint32_t s = lcg_jump(seed, 4u * g);
#pragma unroll
    for (int i = 0; i < 4; i++) idx[i] = (int)(lcg_next(s) % (uint32_t)n);
    return !(idx[0] == idx[1] || idx[0] == idx[2] || idx[0] == idx[3] || idx[1] == idx[2] || idx[1] == idx[3] ||
             idx[2] == idx[3]);
}

// ApplyProjectMat2 (M/matrix.h:1027-1036), reciprocal-multiply form
UAVM_HD void project_mul(float x, float y, const float* M, float& xd, float& yd) {
    float inv = 1.0f / (M[6] * x + M[7] * y + 1.0f);
    xd = (M[0] * x + M[1] * y + M[2]) * inv;
    yd = (M[3] * x + M[4] * y + M[5]) * inv;
}
// ApplyProjectMat3 (M/matrix.h:1002-1012), division form
UAVM_HD void project_div(float x, float y, const float* M, float& xd, float& yd) {
    xd = (M[0] * x + M[1] * y + M[2]) / (M[6] * x + M[7] * y + 1.0f);
    yd = (M[3] * x + M[4] * y + M[5]) / (M[6] * x + M[7] * y + 1.0f);
}
// DistanceSquareOfTwoPoints (M/mvMath.h:209-213)
UAVM_HD float dist2(float xa, float ya, float xb, float yb) {
    return (xa - xb) * (xa - xb) + (ya - yb) * (ya - yb);
}

// ------------------------------------------------------------------------------------------------
// fast path building blocks (4 points -> 8 rows; row 2i = [p0 p1 p2 0 0 0 e f], row 2i+1 = [0 0 0 p0 p1 p2 g h])
// ------------------------------------------------------------------------------------------------
struct Rows4 {
    float p[4][3];   // columns 0-2 of the even row == columns 3-5 of the odd row
    float ef[4][2];  // columns 6,7 of the even row
    float gh[4][2];  // columns 6,7 of the odd row
};

// N = R^T R with MulMatrix's ascending-k accumulation (M/matrix.h:94-120); structurally zero products
// only ever add +-0 and are skipped.
UAVM_HD void normal_matrix(const Rows4& R, float N[8][8]) {
#pragma unroll
    for (int r = 0; r < 3; r++) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            float acc = 0.0f;
#pragma unroll
            for (int i = 0; i < 4; i++) acc += R.p[i][r] * R.p[i][c];
            N[r][c] = acc; N[r + 3][c + 3] = acc;
            N[r][c + 3] = 0.0f; N[r + 3][c] = 0.0f;
        }
#pragma unroll
        for (int c = 0; c < 2; c++) {
            float a0 = 0.0f, a1 = 0.0f;
#pragma unroll
            for (int i = 0; i < 4; i++) { a0 += R.p[i][r] * R.ef[i][c]; a1 += R.p[i][r] * R.gh[i][c]; }
            N[r][6 + c] = a0; N[6 + c][r] = a0;
            N[r + 3][6 + c] = a1; N[6 + c][r + 3] = a1;
        }
    }
#pragma unroll
    for (int r = 0; r < 2; r++)
#pragma unroll
        for (int c = 0; c < 2; c++) {
            float acc = 0.0f;
#pragma unroll
            for (int i = 0; i < 4; i++) { acc += R.ef[i][r] * R.ef[i][c]; acc += R.gh[i][r] * R.gh[i][c]; }
            N[6 + r][6 + c] = acc;
        }
}

// In-place Gauss-Jordan inverse, identical in value to InverseMatrix (M/matrix.h:147-296) when the
// assumptions in the file header hold; returns false otherwise (caller falls back to the slow path).
UAVM_HD bool inverse8_fast(float M[8][8], float eps) {
    bool ok = true;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const float e = M[i][i];
        if (!(fabsf(e) > eps)) ok = false;
#pragma unroll
        for (int c = 0; c < 8; c++)
            if (c != i) M[i][c] = M[i][c] / e;
        M[i][i] = 1.0f / e;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            if (j == i) continue;
            const float f = M[j][i];
            if (fabsf(f) < eps) {
                // The reference skips this row but leaves the non-zero entry f behind in the left block.  For a
                // row that already served as pivot (j < i) that residue is never read again: column i is done,
                // and it can only spread to other rows through a LATER pivot row.  For j > i it can (and could
                // then disturb the exact-1 search of the re-ordering pass), so only that case leaves tier 1.
                if (f != 0.0f && j > i) ok = false;
                M[j][i] = 0.0f;
            } else {
                const float nf = -f;
#pragma unroll
                for (int c = 0; c < 8; c++)
                    if (c != i) M[j][c] = M[j][c] + nf * M[i][c];
                M[j][i] = nf * M[i][i];
            }
        }
    }
    float chk = 0.0f;
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
        for (int c = 0; c < 8; c++) chk += fabsf(M[r][c]);
    if (!(chk <= 3.0e38f)) ok = false;           // inf/NaN anywhere
    return ok;
}

// Tier 2: InverseMatrix (M/matrix.h:147-296) with ALL of its quirks (first-unused-row pivoting, small
// multipliers skipped but not zeroed, row re-ordering by exact-1 search) on the full 8 x 16 augmented
// matrix, written with static indices only so that it stays in registers: the data-dependent pivot /
// swap rows are extracted with select chains instead of dynamic indexing (no local memory).
// Returns 1 ok, 0 when no pivot exists (the reference then leaves its output untouched).
UAVM_HD int inverse8_generic_reg(const float N[8][8], float eps, float out[8][8]) {
    float T[8][16];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int c = 0; c < 8; c++) { T[i][c] = N[i][c]; T[i][8 + c] = (i == c) ? 1.0f : 0.0f; }
    unsigned used = 0u;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        int row = -1;
#pragma unroll
        for (int j = 0; j < 8; j++)
            if (row < 0 && !((used >> j) & 1u) && fabsf(T[j][i]) > eps) row = j;
        if (row < 0) return 0;
        used |= 1u << row;
        float e = 0.0f;
#pragma unroll
        for (int j = 0; j < 8; j++) e = (j == row) ? T[j][i] : e;
        float prow[16];
#pragma unroll
        for (int c = 0; c < 16; c++) {
            float v = 0.0f;
#pragma unroll
            for (int j = 0; j < 8; j++) v = (j == row) ? T[j][c] : v;
            prow[c] = v / e;
        }
#pragma unroll
        for (int j = 0; j < 8; j++) {
            if (j == row) {
#pragma unroll
                for (int c = 0; c < 16; c++) T[j][c] = prow[c];
            } else {
                const float f = T[j][i];
                if (!(fabsf(f) < eps)) {
                    const float nf = -f;
#pragma unroll
                    for (int c = 0; c < 16; c++) T[j][c] = T[j][c] + nf * prow[c];
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < 8; r++) {
        int target = -1;
#pragma unroll
        for (int i = 0; i < 8; i++)
            if (target < 0 && T[i][r] == 1.0f) target = i;
        if (target >= 0 && target != r) {
            float tr[16];
#pragma unroll
            for (int c = 0; c < 16; c++) {
                float v = 0.0f;
#pragma unroll
                for (int i = 0; i < 8; i++) v = (i == target) ? T[i][c] : v;
                tr[c] = v;
            }
#pragma unroll
            for (int i = 0; i < 8; i++)
                if (i == target) {
#pragma unroll
                    for (int c = 0; c < 16; c++) T[i][c] = T[r][c];
                }
#pragma unroll
            for (int c = 0; c < 16; c++) T[r][c] = tr[c];
        }
    }
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int c = 0; c < 8; c++) out[i][c] = T[i][8 + c];
    return 1;
}

// inverse of the 4-point normal matrix of R.  kGeneric == false: tier 1 (in-place fast inverse; false when
// its assumptions fail).  kGeneric == true: tier 2 (register-resident generic inverse); false when no pivot
// exists (the reference then reuses a stale matrix) or the result is not finite (zero-skipping in the
// sparse products would hide NaNs) — those cases go to the tier-3 generic path.
template <bool kGeneric>
UAVM_HD bool normal_inverse(const Rows4& R, float eps, float N[8][8]) {
    if (!kGeneric) {
        normal_matrix(R, N);
        return inverse8_fast(N, eps);
    } else {
        float M[8][8];
        normal_matrix(R, M);
        if (inverse8_generic_reg(M, eps, N) != 1) return false;
        float chk = 0.0f;
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int c = 0; c < 8; c++) chk += fabsf(N[r][c]);
        return chk <= 3.0e38f;
    }
}

// X = (Ninv * R^T) * rhs with the reference association ((A^T A)^-1 A^T) B (M/matrix.h:391-397):
// P[r][k] = sum_m Ninv[r][m] * R[k][m] (ascending m), X[r] = sum_k P[r][k] * rhs[k] (ascending k).
UAVM_HD void solve_from_inverse(const float Ninv[8][8], const Rows4& R, const float rhs[8], float X[8]) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
        float acc = 0.0f;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            float pe = 0.0f, po = 0.0f;
#pragma unroll
            for (int m = 0; m < 3; m++) { pe += Ninv[r][m] * R.p[i][m]; po += Ninv[r][3 + m] * R.p[i][m]; }
#pragma unroll
            for (int m = 0; m < 2; m++) { pe += Ninv[r][6 + m] * R.ef[i][m]; po += Ninv[r][6 + m] * R.gh[i][m]; }
            acc += pe * rhs[2 * i];
            acc += po * rhs[2 * i + 1];
        }
        X[r] = acc;
    }
}

// status of a tuple evaluation
enum { TUPLE_REJECTED = 0, TUPLE_KEPT = 1, TUPLE_REFINED = 2, TUPLE_NEED_SLOW = 3, TUPLE_NEED_REFINE = 4 };

// 4-point direct solve, fast path: SolveHomographyMatrix (M/matrix.h:783-877) + the gates of the sampling
// loop (M/mosaicimage.h:1864-1876).  x1,y1 = points of image 1 (targets), x2,y2 = points of image 2
// (sources).  Returns REJECTED (> 5 px), KEPT (<= 0.01 px, used as is), NEED_REFINE or NEED_SLOW.
template <bool kGeneric>
UAVM_HD int dlt_t(const float x1[4], const float y1[4], const float x2[4], const float y2[4], float h[9]) {
    Rows4 R;
    float rhs[8];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        R.p[i][0] = x2[i]; R.p[i][1] = y2[i]; R.p[i][2] = 1.0f;
        R.ef[i][0] = -x1[i] * x2[i]; R.ef[i][1] = -x1[i] * y2[i];
        R.gh[i][0] = -y1[i] * x2[i]; R.gh[i][1] = -y1[i] * y2[i];
        rhs[2 * i] = x1[i]; rhs[2 * i + 1] = y1[i];
    }
    float N[8][8];
    if (!normal_inverse<kGeneric>(R, 1e-20f, N)) return TUPLE_NEED_SLOW;
    solve_from_inverse(N, R, rhs, h);
    // max residual: float projection, double distance (M/matrix.h:848-866, M/mvMath.h:186-192)
    double e2max = 0.0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        float xf, yf;
        project_mul(x2[i], y2[i], h, xf, yf);
        double dx = (double)x1[i] - (double)xf, dy = (double)y1[i] - (double)yf;
        double e2 = dx * dx + dy * dy;
        if (e2 > e2max) e2max = e2;          // sqrt is monotone: max of sqrt == sqrt of max
    }
    h[8] = (float)sqrt(e2max);
    if (h[8] > 5.0f) return TUPLE_REJECTED;
    if (!((h[8] < 5.0f) && (h[8] > 0.01f))) return TUPLE_KEPT;
    return TUPLE_NEED_REFINE;
}

// Gauss-Newton refine of a 4-point hypothesis, fast path: NonlinearLeastSquareProjection2
// (M/LeastSquare.h:353-531), <= 15 iterations, stop when all |delta| < 1e-10.  h: in = direct solve,
// out = refined (h[8] = max residual, float).  Returns REFINED or NEED_SLOW.
template <bool kGeneric>
UAVM_HD int refine_t(const float x1[4], const float y1[4], const float x2[4], const float y2[4], float h[9]) {
    Rows4 R;
    float N[8][8];
    float w[8];
#pragma unroll
    for (int i = 0; i < 8; i++) w[i] = h[i];
    for (int t = 0; t < 15; t++) {
        float res[8];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const float xs = x2[i], ys = y2[i];
            const float d = w[6] * xs + w[7] * ys + 1.0f;
            const float u = w[0] * xs + w[1] * ys + w[2];
            const float v = w[3] * xs + w[4] * ys + w[5];
            R.p[i][0] = xs / d; R.p[i][1] = ys / d; R.p[i][2] = 1.0f / d;
            const float dd = d * d;
            R.ef[i][0] = -xs * u / dd; R.ef[i][1] = -ys * u / dd;
            R.gh[i][0] = -xs * v / dd; R.gh[i][1] = -ys * v / dd;
            res[2 * i] = x1[i] - u / d;
            res[2 * i + 1] = y1[i] - v / d;
        }
        if (!normal_inverse<kGeneric>(R, 1e-6f, N)) return TUPLE_NEED_SLOW;
        float dx[8];
        solve_from_inverse(N, R, res, dx);
        bool small = true;
#pragma unroll
        for (int i = 0; i < 8; i++) { w[i] += dx[i]; if (!(fabsf(dx[i]) < 1e-10f)) small = false; }
        if (small) break;
    }
#pragma unroll
    for (int i = 0; i < 8; i++) h[i] = w[i];
    float emax = 0.0f;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        float xf, yf;
        project_mul(x2[i], y2[i], h, xf, yf);
        const float ex = x1[i] - xf, ey = y1[i] - yf;
        const float dist = sqrtf(ex * ex + ey * ey);
        if (dist > emax) emax = dist;
    }
    h[8] = emax;
    return TUPLE_REFINED;
}

UAVM_HD int dlt_fast(const float x1[4], const float y1[4], const float x2[4], const float y2[4], float h[9]) {
    return dlt_t<false>(x1, y1, x2, y2, h);
}
UAVM_HD int refine_fast(const float x1[4], const float y1[4], const float x2[4], const float y2[4], float h[9]) {
    return refine_t<false>(x1, y1, x2, y2, h);
}
UAVM_HD int hypothesis_fast(const float x1[4], const float y1[4], const float x2[4], const float y2[4], float h[9]) {
    int st = dlt_fast(x1, y1, x2, y2, h);
    if (st == TUPLE_NEED_REFINE) st = refine_fast(x1, y1, x2, y2, h);
    return st;
}
// tier 2: the whole tuple again with the register-resident generic inverse (rare: ~0.3 % of tuples)
UAVM_HD_NOINLINE int hypothesis_tier2(const float x1[4], const float y1[4], const float x2[4], const float y2[4], float h[9]) {
    int st = dlt_t<true>(x1, y1, x2, y2, h);
    if (st == TUPLE_NEED_REFINE) st = refine_t<true>(x1, y1, x2, y2, h);
    return st;
}

// ------------------------------------------------------------------------------------------------
// slow (generic) path
// ------------------------------------------------------------------------------------------------
// InverseMatrix (M/matrix.h:147-296) for order 8, all quirks kept: first-unused-row pivoting, skipped
// (not zeroed) small multipliers, row re-ordering by exact-1 search.  Returns 1 ok / 0 no pivot.
UAVM_HD_NOINLINE int inverse8_generic(const float* src, float* dst, float eps) {
    const int n = 8, n2 = 16;
    float T[128];
    bool used[8];
    for (int i = 0; i < n; i++) {
        used[i] = false;
        for (int j = 0; j < n2; j++) T[i * n2 + j] = 0.0f;
        T[i * n2 + n + i] = 1.0f;
        for (int j = 0; j < n; j++) T[i * n2 + j] = src[i * n + j];
    }
    for (int i = 0; i < n; i++) {
        float e = 0.0f; int row = 0;
        for (int j = 0; j < n; j++) {
            if (used[j]) continue;
            if (fabsf(T[j * n2 + i]) > eps) { used[j] = true; e = T[j * n2 + i]; row = j; break; }
        }
        if (fabsf(e) < eps) return 0;
        for (int c = 0; c < n2; c++) T[row * n2 + c] = T[row * n2 + c] / e;
        for (int j = 0; j < n; j++) {
            if (j == row) continue;
            if (fabsf(T[j * n2 + i]) < eps) continue;
            const float nf = -T[j * n2 + i];
            for (int c = 0; c < n2; c++) T[j * n2 + c] = T[j * n2 + c] + nf * T[row * n2 + c];
        }
    }
    for (int r = 0; r < n; r++) {
        int target = -1;
        for (int i = 0; i < n && target < 0; i++)
            if (T[i * n2 + r] == 1.0f) target = i;
        if (target >= 0 && target != r)
            for (int j = 0; j < n2; j++) { float t = T[r * n2 + j]; T[r * n2 + j] = T[target * n2 + j]; T[target * n2 + j] = t; }
    }
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) dst[i * n + j] = T[i * n2 + n + j];
    return 1;
}

// dense helpers for the generic path: A is rows x 8 (row-major)
UAVM_HD_NOINLINE void normal_dense(const float* A, int rows, float* N) {
    for (int r = 0; r < 8; r++)
        for (int c = 0; c < 8; c++) {
            float acc = 0.0f;
            for (int k = 0; k < rows; k++) acc += A[k * 8 + r] * A[k * 8 + c];
            N[r * 8 + c] = acc;
        }
}
UAVM_HD_NOINLINE void solve_dense(const float* Ninv, const float* A, int rows, const float* rhs, float* X) {
    for (int r = 0; r < 8; r++) {
        float acc = 0.0f;
        for (int k = 0; k < rows; k++) {
            float p = 0.0f;
            for (int m = 0; m < 8; m++) p += Ninv[r * 8 + m] * A[k * 8 + m];
            acc += p * rhs[k];
        }
        X[r] = acc;
    }
}

UAVM_HD_NOINLINE int hypothesis_slow(const float x1[4], const float y1[4], const float x2[4], const float y2[4], float h[9]) {
    float A[64], rhs[8], N[64], Ninv[64];
    for (int i = 0; i < 64; i++) { A[i] = 0.0f; Ninv[i] = 0.0f; }
    for (int i = 0; i < 4; i++) {
        float* a0 = A + (2 * i) * 8; float* a1 = A + (2 * i + 1) * 8;
        a0[0] = x2[i]; a0[1] = y2[i]; a0[2] = 1.0f; a0[6] = -x1[i] * x2[i]; a0[7] = -x1[i] * y2[i];
        a1[3] = x2[i]; a1[4] = y2[i]; a1[5] = 1.0f; a1[6] = -y1[i] * x2[i]; a1[7] = -y1[i] * y2[i];
        rhs[2 * i] = x1[i]; rhs[2 * i + 1] = y1[i];
    }
    normal_dense(A, 8, N);
    inverse8_generic(N, Ninv, 1e-20f);        // on failure Ninv stays zero (M/matrix.h:360,377)
    solve_dense(Ninv, A, 8, rhs, h);
    double emaxd = 0.0;
    for (int i = 0; i < 4; i++) {
        float xf, yf;
        project_mul(x2[i], y2[i], h, xf, yf);
        double dx = (double)x1[i] - (double)xf, dy = (double)y1[i] - (double)yf;
        double dist = sqrt(dx * dx + dy * dy);
        if (dist > emaxd) emaxd = dist;
    }
    h[8] = (float)emaxd;
    if (h[8] > 5.0f) return TUPLE_REJECTED;
    if (!((h[8] < 5.0f) && (h[8] > 0.01f))) return TUPLE_KEPT;
    float w[8], dx8[8];
    for (int i = 0; i < 8; i++) w[i] = h[i];
    for (int i = 0; i < 64; i++) Ninv[i] = 0.0f;   // reference: uninitialised stack; defined as zeros (oracle.c)
    for (int t = 0; t < 15; t++) {
        for (int i = 0; i < 4; i++) {
            const float xs = x2[i], ys = y2[i];
            const float d = w[6] * xs + w[7] * ys + 1.0f;
            const float u = w[0] * xs + w[1] * ys + w[2];
            const float v = w[3] * xs + w[4] * ys + w[5];
            float* j0 = A + (2 * i) * 8; float* j1 = A + (2 * i + 1) * 8;
            j0[0] = xs / d; j0[1] = ys / d; j0[2] = 1.0f / d; j0[3] = 0.0f; j0[4] = 0.0f; j0[5] = 0.0f;
            j0[6] = -xs * u / (d * d); j0[7] = -ys * u / (d * d);
            j1[0] = 0.0f; j1[1] = 0.0f; j1[2] = 0.0f; j1[3] = xs / d; j1[4] = ys / d; j1[5] = 1.0f / d;
            j1[6] = -xs * v / (d * d); j1[7] = -ys * v / (d * d);
            rhs[2 * i] = x1[i] - u / d;
            rhs[2 * i + 1] = y1[i] - v / d;
        }
        normal_dense(A, 8, N);
        inverse8_generic(N, Ninv, 1e-6f);      // return value ignored (M/LeastSquare.h:451)
        solve_dense(Ninv, A, 8, rhs, dx8);
        bool small = true;
        for (int i = 0; i < 8; i++) { w[i] += dx8[i]; if (!(fabsf(dx8[i]) < 1e-10f)) small = false; }
        if (small) break;
    }
    for (int i = 0; i < 8; i++) h[i] = w[i];
    float emax = 0.0f;
    for (int i = 0; i < 4; i++) {
        float xf, yf;
        project_mul(x2[i], y2[i], h, xf, yf);
        const float ex = x1[i] - xf, ey = y1[i] - yf;
        const float dist = sqrtf(ex * ex + ey * ey);
        if (dist > emax) emax = dist;
    }
    h[8] = emax;
    return TUPLE_REFINED;
}

// one hypothesis: fast path with fallback to the generic path; returns TUPLE_REJECTED/KEPT/REFINED
UAVM_HD int hypothesis(const float x1[4], const float y1[4], const float x2[4], const float y2[4], float h[9],
                       bool* took_slow = nullptr) {
    int st = hypothesis_fast(x1, y1, x2, y2, h);
    if (st == TUPLE_NEED_SLOW) st = hypothesis_tier2(x1, y1, x2, y2, h);
    if (st == TUPLE_NEED_SLOW) {
        if (took_slow) *took_slow = true;
        st = hypothesis_slow(x1, y1, x2, y2, h);
    }
    return st;
}

}  // namespace rmath
}  // namespace uavm
