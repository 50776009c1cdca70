// ransac_warp.cuh — warp-cooperative GENERIC evaluation of one 4-point hypothesis (device only).
//
// The per-thread fast path (ransac_math.cuh) covers ~99.95 % of the tuples.  The rest need the reference's
// InverseMatrix with all of its quirks (M/matrix.h:147-296: first-unused-row pivot search, skipped-but-not-
// zeroed multipliers, row re-ordering by exact-1 search, "return 0 and leave the output untouched").  Doing
// that in one thread needs the full 8 x 16 augmented matrix plus select chains (>255 registers -> local
// memory, ~1 ms per tuple).  Here a whole warp evaluates ONE tuple: lane l owns row l>>2 and the column quad
// l&3 of the augmented matrix (4 floats), pivots are found with a ballot, pivot rows move by shuffles.
// All arithmetic is the dense formulation of the reference (no zero skipping), in its evaluation order, so
// NaN/Inf propagate as they do there, and the stale-inverse quirk of NonlinearLeastSquareProjection2
// (M/LeastSquare.h:451: return value of InverseMatrix ignored) is reproduced by keeping the previous inverse.
#pragma once
#include "ransac_math.cuh"

namespace uavm {
namespace rwarp {

using namespace uavm::rmath;
constexpr unsigned kFull = 0xffffffffu;

// Gauss-Jordan on the distributed augmented matrix.  In: t[0..3] = T[row][4*cq .. 4*cq+3] with the left block
// = matrix to invert and the right block = identity.  Out (return 1): right block = inverse, rows in the
// reference's final order.  Return 0: no pivot (caller keeps its previous inverse).
__device__ __forceinline__ int warp_inverse8(float t[4], float eps, int lane)
{
    const int row = lane >> 2, cq = lane & 3;
    unsigned used = 0u;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        // column i lives in quad (i >> 2), slot (i & 3)
        const float colv = t[i & 3];
        const bool cand = (cq == (i >> 2)) && !((used >> row) & 1u) && (fabsf(colv) > eps);
        const unsigned b = __ballot_sync(kFull, cand);
        if (b == 0u) return 0;                                   // abs(ei) < SMALL_NUMBER -> return 0
        const int plane = __ffs(b) - 1;                          // lowest lane == lowest row index
        const int prow_i = plane >> 2;
        used |= 1u << prow_i;
        const float e = __shfl_sync(kFull, colv, plane);
        const float f = __shfl_sync(kFull, colv, (row << 2) | (i >> 2));     // T[row][i] before this step
        float pr[4];
#pragma unroll
        for (int k = 0; k < 4; k++) pr[k] = __shfl_sync(kFull, t[k], (prow_i << 2) | cq) / e;   // normalised pivot row
        if (row == prow_i) {
#pragma unroll
            for (int k = 0; k < 4; k++) t[k] = pr[k];
        } else if (!(fabsf(f) < eps)) {
            const float nf = -f;
#pragma unroll
            for (int k = 0; k < 4; k++) t[k] = t[k] + nf * pr[k];
        }
    }
    // row re-ordering pass (:244-279): for r = 0..7 find the first row whose column-r entry is exactly 1
#pragma unroll
    for (int r = 0; r < 8; r++) {
        const bool hit = (cq == (r >> 2)) && (t[r & 3] == 1.0f);
        const unsigned b = __ballot_sync(kFull, hit);
        const int target = b ? ((__ffs(b) - 1) >> 2) : -1;
        const bool swap = (target >= 0) && (target != r);        // warp uniform
        const int partner = (row == r) ? target : ((row == target) ? r : row);
        float o[4];
#pragma unroll
        for (int k = 0; k < 4; k++) o[k] = __shfl_sync(kFull, t[k], ((swap ? partner : row) << 2) | cq);
        if (swap) {
#pragma unroll
            for (int k = 0; k < 4; k++) t[k] = o[k];
        }
    }
    return 1;
}

// one dense step shared by the direct solve and the refine: given the 8 x 8 system rows A (every lane holds
// all of A: 8 rows x 8 columns, built redundantly) and the right-hand side, computes
//   N = A^T A (MulMatrix, ascending k), Ninv = InverseMatrix(N, eps) (kept from `ninv_prev` on failure),
//   X = (Ninv A^T) rhs with the reference association.
// ninv[4] holds, for lanes with cq >= 2, Ninv[row][4*(cq-2) .. +3].  Every lane returns all 8 X values.
__device__ __forceinline__ void warp_normal_solve(const float A[8][8], const float rhs[8], float eps, float ninv[4],
                                                  float X[8], int lane)
{
    const int row = lane >> 2, cq = lane & 3;
    float t[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int c = 4 * cq + k;                    // column of the augmented matrix
        float v;
        if (cq < 2) {                                // left block: N[row][c] = sum_k A[k][row] * A[k][c]
            float acc = 0.0f;
#pragma unroll
            for (int kk = 0; kk < 8; kk++) {
                float ar = 0.0f, ac = 0.0f;
#pragma unroll
                for (int j = 0; j < 8; j++) { ar = (j == row) ? A[kk][j] : ar; ac = (j == c) ? A[kk][j] : ac; }
                acc += ar * ac;
            }
            v = acc;
        } else v = (c - 8 == row) ? 1.0f : 0.0f;     // right block: identity
        t[k] = v;
    }
    if (warp_inverse8(t, eps, lane) == 1) {
#pragma unroll
        for (int k = 0; k < 4; k++) ninv[k] = t[k];  // meaningful on lanes with cq >= 2
    }
    // gather Ninv row `row` (8 values) from lanes (row,2) and (row,3)
    float nr[8];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        nr[k] = __shfl_sync(kFull, ninv[k], (row << 2) | 2);
        nr[4 + k] = __shfl_sync(kFull, ninv[k], (row << 2) | 3);
    }
    // X[row] = sum_k (sum_m Ninv[row][m] * A[k][m]) * rhs[k]   (P = Ninv A^T, then P rhs)
    float acc = 0.0f;
#pragma unroll
    for (int kk = 0; kk < 8; kk++) {
        float pk = 0.0f;
#pragma unroll
        for (int m = 0; m < 8; m++) pk += nr[m] * A[kk][m];
        acc += pk * rhs[kk];
    }
#pragma unroll
    for (int r = 0; r < 8; r++) X[r] = __shfl_sync(kFull, acc, r << 2);
}

// Full evaluation of one tuple by a warp (all lanes pass the same points).  Returns TUPLE_REJECTED / KEPT /
// REFINED; h[9] as Ransac2D's aProjectMat2 after the gates (M/mosaicimage.h:1863-1876).
__device__ __noinline__ int warp_hypothesis_generic(const float x1[4], const float y1[4], const float x2[4], const float y2[4],
                                                    float h[9], int lane)
{
    float A[8][8], rhs[8], ninv[4] = {0.0f, 0.0f, 0.0f, 0.0f};       // aATA_Inv starts as zeros (M/matrix.h:360)
#pragma unroll
    for (int i = 0; i < 4; i++) {
#pragma unroll
        for (int c = 0; c < 8; c++) { A[2 * i][c] = 0.0f; A[2 * i + 1][c] = 0.0f; }
        A[2 * i][0] = x2[i]; A[2 * i][1] = y2[i]; A[2 * i][2] = 1.0f; A[2 * i][6] = -x1[i] * x2[i]; A[2 * i][7] = -x1[i] * y2[i];
        A[2 * i + 1][3] = x2[i]; A[2 * i + 1][4] = y2[i]; A[2 * i + 1][5] = 1.0f; A[2 * i + 1][6] = -y1[i] * x2[i]; A[2 * i + 1][7] = -y1[i] * y2[i];
        rhs[2 * i] = x1[i]; rhs[2 * i + 1] = y1[i];
    }
    float X[8];
    warp_normal_solve(A, rhs, 1e-20f, ninv, X, lane);
#pragma unroll
    for (int i = 0; i < 8; i++) h[i] = X[i];
    double emaxd = 0.0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        float xf, yf;
        project_mul(x2[i], y2[i], h, xf, yf);
        const double dx = (double)x1[i] - (double)xf, dy = (double)y1[i] - (double)yf;
        const double dist = sqrt(dx * dx + dy * dy);
        if (dist > emaxd) emaxd = dist;
    }
    h[8] = (float)emaxd;
    if (h[8] > 5.0f) return TUPLE_REJECTED;
    if (!((h[8] < 5.0f) && (h[8] > 0.01f))) return TUPLE_KEPT;

    float w[8];
#pragma unroll
    for (int i = 0; i < 8; i++) w[i] = h[i];
#pragma unroll
    for (int k = 0; k < 4; k++) ninv[k] = 0.0f;      // aJTEMP2: uninitialised in the reference, zeros in the oracle
    for (int it = 0; it < 15; it++) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const float xs = x2[i], ys = y2[i];
            const float d = w[6] * xs + w[7] * ys + 1.0f;
            const float u = w[0] * xs + w[1] * ys + w[2];
            const float v = w[3] * xs + w[4] * ys + w[5];
            A[2 * i][0] = xs / d; A[2 * i][1] = ys / d; A[2 * i][2] = 1.0f / d; A[2 * i][3] = 0.0f; A[2 * i][4] = 0.0f; A[2 * i][5] = 0.0f;
            A[2 * i][6] = -xs * u / (d * d); A[2 * i][7] = -ys * u / (d * d);
            A[2 * i + 1][0] = 0.0f; A[2 * i + 1][1] = 0.0f; A[2 * i + 1][2] = 0.0f; A[2 * i + 1][3] = xs / d; A[2 * i + 1][4] = ys / d; A[2 * i + 1][5] = 1.0f / d;
            A[2 * i + 1][6] = -xs * v / (d * d); A[2 * i + 1][7] = -ys * v / (d * d);
            rhs[2 * i] = x1[i] - u / d;
            rhs[2 * i + 1] = y1[i] - v / d;
        }
        warp_normal_solve(A, rhs, 1e-6f, ninv, X, lane);
        bool small = true;
#pragma unroll
        for (int i = 0; i < 8; i++) { w[i] += X[i]; if (!(fabsf(X[i]) < 1e-10f)) small = false; }
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

}  // namespace rwarp
}  // namespace uavm
