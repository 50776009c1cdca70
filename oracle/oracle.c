/* oracle.c — CPU oracle (TEST INFRASTRUCTURE ONLY; see oracle.h for the pinning status).
 *
 * Plain-C99 restatement of the reference hot path.  "M/" = /root/reference/code/MosaicingCode/
 * mosaicing/.  Every function cites the reference lines it follows.  All float arithmetic is
 * written operation by operation in the reference's evaluation order and compiled with
 * -ffp-contract=off (no FMA), so results are the bit-level truth for the CUDA kernels.
 */
#include "oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------ */
/* sample stream: MSVC rand() (the original runtime of the reference, M/mosaicimage.h:1777)    */
/* ------------------------------------------------------------------------------------------ */
#ifdef _OPENMP
#include <omp.h>
#endif
/* number of host threads used by the parallel loops (torchrun exports OMP_NUM_THREADS=1 by default) */
int orc_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n; return 1;
#endif
}

uint32_t orc_lcg_next(uint32_t* s)
{
    *s = *s * 214013u + 2531011u;
    return (*s >> 16) & 0x7fffu;
}

/* ------------------------------------------------------------------------------------------ */
/* K2: exact brute-force L2 1-NN.  Replaces FlannBasedMatcher().match (M/MosaicWithoutPos.cpp  */
/* :5108-5110).  SIFT descriptors are integers 0..255, so the squared distance is an exact     */
/* integer; ties resolve to the lowest train index.                                             */
/* ------------------------------------------------------------------------------------------ */
void orc_match_l2(const uint8_t* A, int na, const uint8_t* B, int nb, int dim,
                  int32_t* train_idx, int32_t* d2)
{
#pragma omp parallel for schedule(static)
    for (int i = 0; i < na; i++) {
        const uint8_t* a = A + (size_t)i * dim;
        int32_t best = INT32_MAX, bj = -1;
        for (int j = 0; j < nb; j++) {
            const uint8_t* b = B + (size_t)j * dim;
            int32_t acc = 0;
            for (int k = 0; k < dim; k++) {
                int32_t d = (int32_t)a[k] - (int32_t)b[k];
                acc += d * d;
            }
            if (acc < best) { best = acc; bj = j; }
        }
        train_idx[i] = bj;
        d2[i] = best;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* K3: std::sort(matches) (M/MosaicWithoutPos.cpp:5111; DMatch::operator< compares distance     */
/* only, so equal distances are unordered in the reference — we fix the key as (d2, queryIdx)), */
/* nMatch = Min(400, 0.3*size) (:5146-5147), SelectMatchPairs (:4977-5028).                     */
/* Grid quirk (:4993-5010): stepX = width/gridX (int), nX = int(x/stepX) may equal gridX, so    */
/* label[gridX*nY+nX] aliases the next row's first cell or runs past the 9-entry array.  We     */
/* keep the aliasing and give the out-of-range indices (>= gridX*gridY) their own zeroed        */
/* counters (the reference reads heap garbage there).                                           */
/* ------------------------------------------------------------------------------------------ */
typedef struct { int32_t d2; int32_t q; } orc_key;
static int orc_key_cmp(const void* a, const void* b)
{
    const orc_key* x = (const orc_key*)a; const orc_key* y = (const orc_key*)b;
    if (x->d2 != y->d2) return x->d2 < y->d2 ? -1 : 1;
    return x->q < y->q ? -1 : (x->q > y->q ? 1 : 0);
}
int orc_select(const int32_t* train_idx, const int32_t* d2, int n,
               const float* kp1_xy, const float* kp2_xy,
               int width, int height, int grid_x, int grid_y, int max_num, double frac,
               float* out_xy1, int32_t* out_id1, float* out_xy2, int32_t* out_id2)
{
    if (n <= 0) return 0;
    orc_key* keys = (orc_key*)malloc(sizeof(orc_key) * (size_t)n);
    for (int i = 0; i < n; i++) { keys[i].d2 = d2[i]; keys[i].q = i; }
    qsort(keys, (size_t)n, sizeof(orc_key), orc_key_cmp);

    double nm_d = frac * (double)n;                       /* 0.3*matches.size() */
    int n_match = (int)(((double)max_num < nm_d) ? (double)max_num : nm_d);   /* Min macro, then int */
    int n_grids = grid_x * grid_y;
    int per_grid = (int)((float)n_match / n_grids);       /* :4989 */
    int step_x = width / grid_x, step_y = height / grid_y;
    int n_label = grid_x * (grid_y + 1) + grid_x + 1;
    int* label = (int*)calloc((size_t)n_label, sizeof(int));
    int cnt = 0;
    for (int k = 0; k < n; k++) {
        int q = keys[k].q;
        float x = kp1_xy[2 * q], y = kp1_xy[2 * q + 1];
        int nx = (int)(x / step_x);                      /* float / int -> float, truncation */
        int ny = (int)(y / step_y);
        int cell = grid_x * ny + nx;
        if (cell < 0 || cell >= n_label) continue;       /* negative coords: outside any counter */
        if (label[cell] >= per_grid) continue;
        int t = train_idx[q];
        out_xy1[2 * cnt] = x; out_xy1[2 * cnt + 1] = y; out_id1[cnt] = q;
        out_xy2[2 * cnt] = kp2_xy[2 * t]; out_xy2[2 * cnt + 1] = kp2_xy[2 * t + 1]; out_id2[cnt] = t;
        cnt++;
        label[cell]++;
    }
    free(label);
    free(keys);
    return cnt;
}

/* ------------------------------------------------------------------------------------------ */
/* K4 numerics                                                                                  */
/* ------------------------------------------------------------------------------------------ */

/* MulMatrix (M/matrix.h:94-120): naive triple loop, ascending inner index, float accumulate. */
static void mul_matrix(const float* a, int r1, int c1, const float* b, int c2, float* d)
{
    for (int r = 0; r < r1; r++)
        for (int c = 0; c < c2; c++) {
            float acc = 0;
            for (int k = 0; k < c1; k++) acc += a[r * c1 + k] * b[k * c2 + c];
            d[r * c2 + c] = acc;
        }
}
/* TransposeMatrix (M/matrix.h:70-90) */
static void transpose_matrix(const float* s, int row, int col, float* d)
{
    for (int r = 0; r < row; r++)
        for (int c = 0; c < col; c++) d[c * row + r] = s[r * col + c];
}

/* InverseMatrix (M/matrix.h:147-296): Gauss-Jordan on the augmented [S | I]; pivot = FIRST unused
 * row whose entry exceeds eps (no partial pivoting); rows whose entry is below eps are skipped,
 * not zeroed; final pass re-orders rows by looking for exact 1.0 entries.  Returns 1 ok, 0 if no
 * pivot was found (dst untouched), -1 bad order. */
int orc_inverse_matrix(const float* src, int order, float* dst, float eps)
{
    if (order > 13 || order < 2) return -1;
    float T[400];
    int used[16];
    int o2 = order * 2;
    memset(T, 0, sizeof(float) * (size_t)(order * o2));
    for (int i = 0; i < order; i++) {
        used[i] = 0;
        T[i * o2 + order + i] = 1;
        for (int j = 0; j < order; j++) T[i * o2 + j] = src[i * order + j];
    }
    for (int i = 0; i < order; i++) {            /* column */
        float ei = 0; int row_i = 0;
        for (int j = 0; j < order; j++) {
            if (used[j]) continue;
            if (fabsf(T[j * o2 + i]) > eps) { used[j] = 1; ei = T[j * o2 + i]; row_i = j; break; }
        }
        if (fabsf(ei) < eps) return 0;
        for (int c = 0; c < o2; c++) T[row_i * o2 + c] /= ei;
        for (int j = 0; j < order; j++) {
            if (j == row_i) continue;
            if (fabsf(T[j * o2 + i]) < eps) continue;
            float e2 = T[j * o2 + i];
            for (int c = 0; c < o2; c++) T[j * o2 + c] += -e2 * T[row_i * o2 + c];
        }
    }
    for (int r = 0; r < order; r++) {            /* row re-ordering pass (:244-279) */
        int target_row = -1;
        for (int i = 0; i < order && target_row < 0; i++)
            for (int j = 0; j < order; j++)
                if (T[i * o2 + j] == 1 && j == r) { target_row = i; break; }
        if (target_row >= 0 && target_row != r)
            for (int j = 0; j < o2; j++) {
                float t = T[r * o2 + j]; T[r * o2 + j] = T[target_row * o2 + j]; T[target_row * o2 + j] = t;
            }
    }
    for (int i = 0; i < order; i++)
        for (int j = 0; j < order; j++) dst[i * order + j] = T[i * o2 + order + j];
    return 1;
}

/* ApplyProjectMat2 (M/matrix.h:1027-1036): reciprocal-multiply form, float. */
static void apply_project_mat2(float x, float y, float* xd, float* yd, const float* M)
{
    float inv = 1 / (M[6] * x + M[7] * y + 1);
    *xd = (M[0] * x + M[1] * y + M[2]) * inv;
    *yd = (M[3] * x + M[4] * y + M[5]) * inv;
}
/* ApplyProjectMat3 (M/matrix.h:1002-1012): division form, float. */
static void apply_project_mat3(float x, float y, float* xd, float* yd, const float* M)
{
    *xd = (M[0] * x + M[1] * y + M[2]) / (M[6] * x + M[7] * y + 1);
    *yd = (M[3] * x + M[4] * y + M[5]) / (M[6] * x + M[7] * y + 1);
}
/* DistanceSquareOfTwoPoints (M/mvMath.h:209-213) */
static float dist2f(float x1, float y1, float x2, float y2)
{
    return (x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2);
}

/* SolveLinearLeastSquare2 (M/matrix.h:334-403): X = ((A^T A)^-1 A^T) B, float, eps 1e-20. */
static void solve_lls2(const float* A, int rowA, const float* B, float* X)
{
    const int colA = 8;
    float* AT = (float*)malloc(sizeof(float) * (size_t)(rowA * colA));
    float* P = (float*)malloc(sizeof(float) * (size_t)(rowA * colA));
    transpose_matrix(A, rowA, colA, AT);
    float ATA[64], ATAinv[64];
    memset(ATAinv, 0, sizeof(ATAinv));                   /* T1 aATA_Inv[20*20] = {0} (:360) */
    mul_matrix(AT, colA, rowA, A, colA, ATA);
    orc_inverse_matrix(ATA, colA, ATAinv, 1e-20f);       /* return value ignored (:377) */
    mul_matrix(ATAinv, colA, colA, AT, rowA, P);
    mul_matrix(P, colA, rowA, B, 1, X);
    free(AT); free(P);
}

/* SolveHomographyMatrix (M/matrix.h:783-877): maps points2 -> points1; h[8] = max residual,
 * evaluated with a float projection but a DOUBLE distance (:848-866, M/mvMath.h:186-192). */
int orc_solve_homography(const float* xy1, const float* xy2, int n, float h[9])
{
    if (n < 4) return 0;
    int rowA = 2 * n;
    float* A = (float*)calloc((size_t)(rowA * 8), sizeof(float));
    float* B = (float*)calloc((size_t)rowA, sizeof(float));
    for (int r = 0; r < n; r++) {
        float x1 = xy1[2 * r], y1 = xy1[2 * r + 1], x2 = xy2[2 * r], y2 = xy2[2 * r + 1];
        float* a0 = A + (2 * r) * 8; float* a1 = A + (2 * r + 1) * 8;
        a0[0] = x2; a0[1] = y2; a0[2] = 1; a0[6] = -x1 * x2; a0[7] = -x1 * y2;
        a1[3] = x2; a1[4] = y2; a1[5] = 1; a1[6] = -y1 * x2; a1[7] = -y1 * y2;
        B[2 * r] = x1; B[2 * r + 1] = y1;
    }
    float X[9];
    solve_lls2(A, rowA, B, X);
    for (int i = 0; i < 8; i++) h[i] = X[i];
    double err_max = 0;
    for (int i = 0; i < n; i++) {
        float xf, yf;
        apply_project_mat2(xy2[2 * i], xy2[2 * i + 1], &xf, &yf, h);
        double xd = xf, yd = yf;
        double dx = (double)xy1[2 * i] - xd, dy = (double)xy1[2 * i + 1] - yd;
        double dist = sqrt(dx * dx + dy * dy);
        if (dist > err_max) err_max = dist;
    }
    h[8] = (float)err_max;
    free(A); free(B);
    return 1;
}

/* NonlinearLeastSquareProjection2 (M/LeastSquare.h:353-531): <=15 Gauss-Newton iterations,
 * InverseMatrix with the default eps 1e-6 whose return value is ignored (:451): on failure the
 * previous iteration's inverse is reused (first iteration: the reference's array is
 * uninitialised stack memory; the oracle defines it as zeros and counts the event). */
int orc_nls_projection2(const float* xy1, const float* xy2, int n, float out[9], const float init[9],
                        float stop, int* n_inv_fail)
{
    if (n < 4) return 0;
    const int D = 8;
    int rows = 2 * n;
    float* J = (float*)malloc(sizeof(float) * (size_t)(rows * D));
    float* C = (float*)malloc(sizeof(float) * (size_t)rows);
    float* JT = (float*)malloc(sizeof(float) * (size_t)(rows * D));
    float* JL = (float*)malloc(sizeof(float) * (size_t)(rows * D));
    float N1[64], N2[64], dx[8], w[9];
    memset(N2, 0, sizeof(N2));
    w[8] = 0;
    for (int t = 0; t < 15; t++) {
        if (t == 0) memcpy(w, init, sizeof(float) * 8);
        for (int i = 0; i < n; i++) {
            float x2 = xy1[2 * i], y2 = xy1[2 * i + 1];     /* matchsort1 = target */
            float x1 = xy2[2 * i], y1 = xy2[2 * i + 1];     /* matchsort0 = source */
            float d = w[6] * x1 + w[7] * y1 + 1;
            float u = w[0] * x1 + w[1] * y1 + w[2];
            float v = w[3] * x1 + w[4] * y1 + w[5];
            float* j0 = J + i * 16;
            j0[0] = x1 / d; j0[1] = y1 / d; j0[2] = 1 / d;
            j0[3] = 0; j0[4] = 0; j0[5] = 0;
            j0[6] = -x1 * u / (d * d); j0[7] = -y1 * u / (d * d);
            j0[8] = 0; j0[9] = 0; j0[10] = 0;
            j0[11] = x1 / d; j0[12] = y1 / d; j0[13] = 1 / d;
            j0[14] = -x1 * v / (d * d); j0[15] = -y1 * v / (d * d);
            C[2 * i] = x2 - u / d;
            C[2 * i + 1] = y2 - v / d;
        }
        transpose_matrix(J, rows, D, JT);
        mul_matrix(JT, D, rows, J, D, N1);
        if (orc_inverse_matrix(N1, D, N2, 1e-6f) != 1 && n_inv_fail) (*n_inv_fail)++;
        mul_matrix(N2, D, D, JT, rows, JL);
        mul_matrix(JL, D, rows, C, 1, dx);
        for (int i = 0; i < 8; i++) w[i] += dx[i];
        int all_small = 1;
        for (int i = 0; i < 8; i++) if (!(fabsf(dx[i]) < stop)) all_small = 0;
        if (all_small) break;
    }
    for (int i = 0; i < 8; i++) out[i] = w[i];
    float err_max = 0;
    for (int i = 0; i < n; i++) {
        float xd, yd;
        apply_project_mat2(xy2[2 * i], xy2[2 * i + 1], &xd, &yd, out);
        float ex = xy1[2 * i] - xd, ey = xy1[2 * i + 1] - yd;
        float dist = sqrtf(ex * ex + ey * ey);
        if (dist > err_max) err_max = dist;
    }
    out[8] = err_max;
    free(J); free(C); free(JT); free(JL);
    return 1;
}

/* loop body of Ransac2D for one 4-tuple (M/mosaicimage.h:1855-1904) */
int orc_ransac_eval_tuple(const float* xy1, const float* xy2, int n, const int32_t idx[4],
                          float thr2, float h[9], int* support)
{
    float s1[8], s2[8];
    for (int i = 0; i < 4; i++) {
        s1[2 * i] = xy1[2 * idx[i]]; s1[2 * i + 1] = xy1[2 * idx[i] + 1];
        s2[2 * i] = xy2[2 * idx[i]]; s2[2 * i + 1] = xy2[2 * idx[i] + 1];
    }
    for (int i = 0; i < 9; i++) h[i] = 0;
    int kind = 1;
    orc_solve_homography(s1, s2, 4, h);
    if (h[8] > 5) { if (support) *support = 0; return 0; }
    else if ((h[8] < 5) && (h[8] > 0.01f)) {
        float fine[9];
        orc_nls_projection2(s1, s2, 4, fine, h, 1e-10f, 0);
        memcpy(h, fine, sizeof(fine));
        kind = 2;
    }
    int cur = 0;
    for (int i = 0; i < n; i++) {
        float xb, yb;
        apply_project_mat2(xy2[2 * i], xy2[2 * i + 1], &xb, &yb, h);
        if (dist2f(xb, yb, xy1[2 * i], xy1[2 * i + 1]) < thr2) cur++;
    }
    if (support) *support = cur;
    return kind;
}

/* Ransac2D (M/mosaicimage.h:1729-2035), PROJECT_MODEL branch. */
int orc_ransac2d(const float* xy1, const float* xy2, int n, float ransac_dist, int sample_times,
                 uint32_t seed, uint8_t* inlier_mask, float H[9], int* n_inliers,
                 orc_ransac_stats* stats)
{
    orc_ransac_stats st; memset(&st, 0, sizeof(st)); st.best_tuple = -1;
    if (inlier_mask && n > 0) memset(inlier_mask, 0, (size_t)n);
    if (n_inliers) *n_inliers = 0;
    for (int i = 0; i < 9; i++) H[i] = 0;
    if (stats) *stats = st;
    if (n <= 0) return 0;
    float thr2 = ransac_dist * ransac_dist;
    if (n < 4) return 0;
    float inv_n = 1.0f / n;
    const int max_times = 5000;
    if (sample_times > max_times) sample_times = max_times;
    float* mats = (float*)calloc((size_t)(sample_times > 0 ? sample_times : 1) * 9, sizeof(float));
    uint32_t state = seed;                               /* srand(seed) */
    int max_support = 0, max_idx = 0, real = 0;
    for (int t = 0; t < sample_times;) {
        real++;
        if (real >= max_times) break;
        int32_t idx[4];
        do {
            for (int i = 0; i < 4; i++) { idx[i] = (int32_t)(orc_lcg_next(&state) % (uint32_t)n); st.rand_calls++; }
        } while (idx[0] == idx[1] || idx[0] == idx[2] || idx[0] == idx[3] ||
                 idx[1] == idx[2] || idx[1] == idx[3] || idx[2] == idx[3]);
        st.n_tuples = real;
        float h[9]; int sup = 0;
        float s1[8], s2[8];
        for (int i = 0; i < 4; i++) {
            s1[2 * i] = xy1[2 * idx[i]]; s1[2 * i + 1] = xy1[2 * idx[i] + 1];
            s2[2 * i] = xy2[2 * idx[i]]; s2[2 * i + 1] = xy2[2 * idx[i] + 1];
        }
        for (int i = 0; i < 9; i++) h[i] = 0;
        orc_solve_homography(s1, s2, 4, h);
        if (h[8] > 5) continue;
        else if ((h[8] < 5) && (h[8] > 0.01f)) {
            float fine[9];
            orc_nls_projection2(s1, s2, 4, fine, h, 1e-10f, &st.n_inv_fail);
            memcpy(h, fine, sizeof(fine));
            st.n_refined++;
        }
        memcpy(mats + (size_t)t * 9, h, sizeof(float) * 9);
        for (int i = 0; i < n; i++) {
            float xb, yb;
            apply_project_mat2(xy2[2 * i], xy2[2 * i + 1], &xb, &yb, h);
            if (dist2f(xb, yb, xy1[2 * i], xy1[2 * i + 1]) < thr2) sup++;
        }
        if (sup > max_support) {
            max_support = sup; max_idx = t; st.best_tuple = real - 1;
            if (max_support * inv_n > 0.99f) { st.early_exit = 1; t++; break; }
        }
        t++;
        st.n_counted = t;
    }
    if (st.early_exit) st.n_counted = max_idx + 1;
    st.best_t = max_idx; st.max_support = max_support;
    const float* best = mats + (size_t)max_idx * 9;      /* zeros if nothing was ever stored */
    int cnt = 0;
    float* in1 = (float*)malloc(sizeof(float) * 2 * (size_t)n);
    float* in2 = (float*)malloc(sizeof(float) * 2 * (size_t)n);
    for (int i = 0; i < n; i++) {
        float xb, yb;
        apply_project_mat3(xy2[2 * i], xy2[2 * i + 1], &xb, &yb, best);
        if (dist2f(xb, yb, xy1[2 * i], xy1[2 * i + 1]) < thr2) {
            if (inlier_mask) inlier_mask[i] = 1;
            in1[2 * cnt] = xy1[2 * i]; in1[2 * cnt + 1] = xy1[2 * i + 1];
            in2[2 * cnt] = xy2[2 * i]; in2[2 * cnt + 1] = xy2[2 * i + 1];
            cnt++;
        }
    }
    int success = 1;
    if (cnt > 0) { if (!orc_solve_homography(in1, in2, cnt, H)) success = 0; }
    else success = 0;
    if (success) {
        float motion[9];
        orc_nls_projection2(in1, in2, cnt, motion, best, 1e-10f, &st.n_inv_fail);
        memcpy(H, motion, sizeof(motion));
    }
    if (n_inliers) *n_inliers = cnt;
    if (stats) *stats = st;
    free(in1); free(in2); free(mats);
    return cnt >= 4 ? 1 : 0;
}

/* ------------------------------------------------------------------------------------------ */
/* K8: BundleAdjustmentSparse (M/MosaicWithoutPos.cpp:6971-7202) + SolveSparseSystem2           */
/* (M/test_cholmod.cpp:180-262: x = (A^T A)^-1 A^T b by sparse Cholesky, double).  The oracle   */
/* assembles the normal equations densely in double and solves with a dense Cholesky (LL^T).    */
/* ------------------------------------------------------------------------------------------ */
static int cholesky_solve(double* N, double* b, int n)
{
    for (int j = 0; j < n; j++) {
        double s = N[j * n + j];
        for (int k = 0; k < j; k++) s -= N[j * n + k] * N[j * n + k];
        if (!(s > 0)) return -1;
        double l = sqrt(s);
        N[j * n + j] = l;
        for (int i = j + 1; i < n; i++) {
            double t = N[i * n + j];
            for (int k = 0; k < j; k++) t -= N[i * n + k] * N[j * n + k];
            N[i * n + j] = t / l;
        }
    }
    for (int i = 0; i < n; i++) {
        double t = b[i];
        for (int k = 0; k < i; k++) t -= N[i * n + k] * b[k];
        b[i] = t / N[i * n + i];
    }
    for (int i = n - 1; i >= 0; i--) {
        double t = b[i];
        for (int k = i + 1; k < n; k++) t -= N[k * n + i] * b[k];
        b[i] = t / N[i * n + i];
    }
    return 0;
}
int orc_align_affine(const double* pairs, int n_pairs, const int32_t* fixed_img, const float* T0,
                     int n_images, float* out)
{
    if (n_images <= 1) return -2;
    int* acc_fixed = (int*)calloc((size_t)n_images, sizeof(int));
    int n_fixed = 0;
    for (int i = 0; i < n_images; i++) {
        acc_fixed[i] = n_fixed;                          /* #fixed images before i (:6985-6997) */
        if (fixed_img[i] == 1) n_fixed++;
    }
    int nu = 6 * (n_images - n_fixed);
    if (nu <= 0) { free(acc_fixed); return -2; }
    double* N = (double*)calloc((size_t)nu * nu, sizeof(double));
    double* g = (double*)calloc((size_t)nu, sizeof(double));
    /* column slots inside an image block: 0:a 1:b 4:e for the x-row, 2:c 3:d 5:f for the y-row */
    static const int slot_x[3] = {0, 1, 4}, slot_y[3] = {2, 3, 5};
    for (int n = 0; n < n_pairs; n++) {
        const double* p = pairs + (size_t)n * 8;
        int ia = (int)p[0], fa = (int)p[3], ib = (int)p[4], fb = (int)p[7];
        double xa = p[1], ya = p[2], xb = p[5], yb = p[6];
        int cols[6]; double vals[6]; int nc = 0; double rhs_x = 0, rhs_y = 0;
        int use = 1;
        if (fa == 0 && fb == 0) {
            int ca = 6 * (ia - acc_fixed[ia]), cb = 6 * (ib - acc_fixed[ib]);
            cols[0] = ca; vals[0] = xa; cols[1] = ca; vals[1] = ya; cols[2] = ca; vals[2] = 1;
            cols[3] = cb; vals[3] = -xb; cols[4] = cb; vals[4] = -yb; cols[5] = cb; vals[5] = -1;
            nc = 6;
        } else if (fa == 1 && fb == 0) {
            int cb = 6 * (ib - acc_fixed[ib]);
            cols[0] = cb; vals[0] = xb; cols[1] = cb; vals[1] = yb; cols[2] = cb; vals[2] = 1; nc = 3;
            double h[9]; for (int t = 0; t < 9; t++) h[t] = T0[(size_t)ia * 9 + t];
            rhs_x = (h[0] * xa + h[1] * ya + h[2]) / (h[6] * xa + h[7] * ya + h[8]);
            rhs_y = (h[3] * xa + h[4] * ya + h[5]) / (h[6] * xa + h[7] * ya + h[8]);
        } else if (fa == 0 && fb == 1) {
            int ca = 6 * (ia - acc_fixed[ia]);
            cols[0] = ca; vals[0] = xa; cols[1] = ca; vals[1] = ya; cols[2] = ca; vals[2] = 1; nc = 3;
            double h[9]; for (int t = 0; t < 9; t++) h[t] = T0[(size_t)ib * 9 + t];
            rhs_x = (h[0] * xb + h[1] * yb + h[2]) / (h[6] * xb + h[7] * yb + h[8]);
            rhs_y = (h[3] * xb + h[4] * yb + h[5]) / (h[6] * xb + h[7] * yb + h[8]);
        } else use = 0;
        if (!use) continue;
        /* x-row uses slots {a,b,e}, y-row uses slots {c,d,f} with the same values */
        for (int a = 0; a < nc; a++) {
            int ca_x = cols[a] + slot_x[a % 3], ca_y = cols[a] + slot_y[a % 3];
            for (int b = 0; b < nc; b++) {
                int cb_x = cols[b] + slot_x[b % 3], cb_y = cols[b] + slot_y[b % 3];
                N[(size_t)ca_x * nu + cb_x] += vals[a] * vals[b];
                N[(size_t)ca_y * nu + cb_y] += vals[a] * vals[b];
            }
            g[ca_x] += vals[a] * rhs_x;
            g[ca_y] += vals[a] * rhs_y;
        }
    }
    int rc = cholesky_solve(N, g, nu);
    if (rc == 0) {
        int k = 0;
        for (int i = 0; i < n_images; i++) {
            float* o = out + (size_t)i * 9;
            if (fixed_img[i] == 0) {
                o[0] = (float)g[6 * k + 0]; o[1] = (float)g[6 * k + 1];
                o[3] = (float)g[6 * k + 2]; o[4] = (float)g[6 * k + 3];
                o[2] = (float)g[6 * k + 4]; o[5] = (float)g[6 * k + 5];
                o[6] = 0; o[7] = 0; o[8] = 1;
                k++;
            } else memcpy(o, T0 + (size_t)i * 9, sizeof(float) * 9);
        }
    }
    free(N); free(g); free(acc_fixed);
    return rc;
}

/* ------------------------------------------------------------------------------------------ */
/* K5: canvas layout (M/MosaicImage.cpp:2233-2348) and chip warp (:2350-2448)                   */
/* ------------------------------------------------------------------------------------------ */
int orc_canvas_layout_compute(const float* H, const int32_t* keep_in, int n, int img_w, int img_h,
                              orc_canvas_layout* canvas, orc_chip_layout* chips)
{
    float maxX = 0, maxY = 0, minX = 0, minY = 0;
    float* beg = (float*)calloc((size_t)n * 2, sizeof(float));
    float* end = (float*)calloc((size_t)n * 2, sizeof(float));
    float cx[4] = {0, (float)(img_w - 1), (float)(img_w - 1), 0};
    float cy[4] = {0, 0, (float)(img_h - 1), (float)(img_h - 1)};
    for (int k = 0; k < n; k++) {
        const float* m = H + (size_t)k * 9;
        chips[k].keep = 0;
        if (keep_in && keep_in[k] == 0) continue;
        if (m[8] == 0) continue;
        chips[k].keep = 1;
        float bmaxx = (float)(-(1 << 29)), bmaxy = (float)(-(1 << 29)), bminx = (float)(1 << 29), bminy = (float)(1 << 29);
        for (int i = 0; i < 4; i++) {
            float xs = cx[i], ys = cy[i];
            float xd = (xs * m[0] + ys * m[1] + m[2]) / (xs * m[6] + ys * m[7] + m[8]);
            float yd = (xs * m[3] + ys * m[4] + m[5]) / (xs * m[6] + ys * m[7] + m[8]);
            if (xd > maxX) maxX = xd;
            if (xd < minX) minX = xd;
            if (yd > maxY) maxY = yd;
            if (yd < minY) minY = yd;
            if (xd > bmaxx) bmaxx = xd;
            if (xd < bminx) bminx = xd;
            if (yd > bmaxy) bmaxy = yd;
            if (yd < bminy) bminy = yd;
        }
        beg[2 * k] = bminx; beg[2 * k + 1] = bminy; end[2 * k] = bmaxx; end[2 * k + 1] = bmaxy;
    }
    float dgx = -minX, dgy = -minY;
    canvas->dgx = dgx; canvas->dgy = dgy;
    canvas->canvas_w = (int)(maxX - minX + 1.5f);
    canvas->canvas_h = (int)(maxY - minY + 1.5f);
    for (int k = 0; k < n; k++) {
        if (!chips[k].keep) continue;
        const float* m = H + (size_t)k * 9;
        float bx = beg[2 * k] + dgx, by = beg[2 * k + 1] + dgy;
        float ex = end[2 * k] + dgx, ey = end[2 * k + 1] + dgy;
        int ibx = (int)bx, iby = (int)by, iex = (int)(ex + 0.5f), iey = (int)(ey + 0.5f);
        float sx = ibx - bx, sy = iby - by;
        chips[k].beg_x = ibx; chips[k].beg_y = iby;
        chips[k].chip_w = iex - ibx + 1; chips[k].chip_h = iey - iby + 1;
        chips[k].sx = sx; chips[k].sy = sy;
        for (int i = 0; i < 4; i++) {
            /* ApplyProjectMat9 (M/matrix.h:1015-1024): inv = 1/(m6 x + m7 y + m8), multiply */
            float inv = 1 / (m[6] * cx[i] + m[7] * cy[i] + m[8]);
            float tx = (m[0] * cx[i] + m[1] * cy[i] + m[2]) * inv;
            float ty = (m[3] * cx[i] + m[4] * cy[i] + m[5]) * inv;
            chips[k].quad[2 * i] = tx + dgx + sx - ibx;
            chips[k].quad[2 * i + 1] = ty + dgy + sy - iby;
        }
        memset(chips[k].inv, 0, sizeof(chips[k].inv));
        orc_inverse_matrix(m, 3, chips[k].inv, 1e-12f);
    }
    free(beg); free(end);
    return 0;
}

void orc_warp_chip(const uint8_t* src, int img_w, int img_h, int src_step,
                   const orc_canvas_layout* canvas, const orc_chip_layout* chip,
                   uint8_t* chip_px, int chip_step, uint8_t* mask, int mask_step)
{
    const float* iv = chip->inv;
    float dgx = canvas->dgx, dgy = canvas->dgy, sx = chip->sx, sy = chip->sy;
    int bx = chip->beg_x, by = chip->beg_y;
    int w1 = img_w - 1, h1 = img_h - 1;
#pragma omp parallel for schedule(static)
    for (int yd = 0; yd < chip->chip_h; yd++) {
        uint8_t* row = chip_px + (size_t)yd * chip_step;
        uint8_t* mrow = mask + (size_t)yd * mask_step;
        for (int xd = 0; xd < chip->chip_w; xd++) {
            float xt = xd - dgx - sx + bx;
            float yt = yd - dgy - sy + by;
            float xs = (xt * iv[0] + yt * iv[1] + iv[2]) / (xt * iv[6] + yt * iv[7] + iv[8]);
            float ys = (xt * iv[3] + yt * iv[4] + iv[5]) / (xt * iv[6] + yt * iv[7] + iv[8]);
            int iy = (int)ys, ix = (int)xs;
            if ((xs >= 0) && (xs < w1) && (ys >= 0) && (ys < h1)) {
                float p = ys - iy, q = xs - ix;
                const uint8_t* t = src + (size_t)iy * src_step + 3 * ix;
                for (int c = 0; c < 3; c++) {
                    int g1 = t[c], g2 = t[c + 3], g3 = t[c + src_step], g4 = t[c + src_step + 3];
                    float v = g1 * (1 - p) * (1 - q) + g2 * (1 - p) * q + g3 * p * (1 - q) + g4 * p * q;
                    row[3 * xd + c] = (uint8_t)(int)v;
                }
                mrow[xd] = 255;
            } else {
                row[3 * xd] = 0; row[3 * xd + 1] = 0; row[3 * xd + 2] = 0;
                mrow[xd] = 0;
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* K6: FindMasksByDistMap (M/MosaicImage.cpp:1761-1881); LineOf2Points1 (M/ImageMath.cpp:88-103) */
/* ------------------------------------------------------------------------------------------ */
static void line_of_2_points(float* a, float* b, float* c, float x1, float y1, float x2, float y2)
{
    if (fabs(x1 - x2) < 0.000001) { *a = 1.0f; *b = 0; *c = -x1; }
    else { *a = (y1 - y2) / (x1 - x2); *b = -1.0f; *c = y1 - *a * x1; }
}
int orc_seam_masks(uint8_t** masks, const int32_t* mask_step, const orc_chip_layout* chips,
                   int n_valid, int canvas_w, int canvas_h)
{
    float** maps = (float**)calloc((size_t)n_valid, sizeof(float*));
    for (int n = 0; n < n_valid; n++) {
        const float* q = chips[n].quad;
        float A[4], B[4], C[4], inv[4];
        for (int i = 0; i < 4; i++) {
            int j = (i + 1) & 3;
            line_of_2_points(&A[i], &B[i], &C[i], q[2 * i], q[2 * i + 1], q[2 * j], q[2 * j + 1]);
            inv[i] = 1 / sqrtf(A[i] * A[i] + B[i] * B[i]);
        }
        int w = chips[n].chip_w, h = chips[n].chip_h, ws = mask_step[n];
        float* map = (float*)calloc((size_t)ws * h, sizeof(float));
        float max_dist = 0;
        for (int r = 0; r < h; r++) {
            const uint8_t* mrow = masks[n] + (size_t)ws * r;
            float* prow = map + (size_t)ws * r;
            for (int c = 0; c < w; c++) {
                if (mrow[c] == 0) continue;
                float mind = (float)(1 << 29);
                for (int i = 0; i < 4; i++) {
                    float d = fabsf(A[i] * c + B[i] * r + C[i]) * inv[i];
                    if (d < mind) mind = d;
                }
                prow[c] = mind;
                if (mind > max_dist) max_dist = mind;
            }
        }
        for (int r = 0; r < h; r++) {
            float* prow = map + (size_t)ws * r;
            for (int c = 0; c < w; c++) prow[c] /= max_dist;
        }
        maps[n] = map;
    }
    for (int n = 0; n < n_valid; n++) memset(masks[n], 0, (size_t)mask_step[n] * chips[n].chip_h);
    for (int r = 0; r < canvas_h; r++)
        for (int c = 0; c < canvas_w; c++) {
            int max_idx = -1; float max_dist = 0;
            for (int n = 0; n < n_valid; n++) {
                int yc = r - chips[n].beg_y, xc = c - chips[n].beg_x;
                if (yc >= 0 && yc < chips[n].chip_h && xc >= 0 && xc < chips[n].chip_w) {
                    float cur = maps[n][(size_t)yc * mask_step[n] + xc];
                    if (cur > max_dist) { max_dist = cur; max_idx = n; }
                }
            }
            if (max_idx >= 0)
                masks[max_idx][(size_t)mask_step[max_idx] * (r - chips[max_idx].beg_y) + (c - chips[max_idx].beg_x)] = 255;
        }
    for (int n = 0; n < n_valid; n++) free(maps[n]);
    free(maps);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* MosaicImagesRefined (M/MosaicWithoutPos.cpp:2194-2352): no blending, images pasted in index order */
/* ------------------------------------------------------------------------------------------ */
static void apply_project9(const float* h, float x, float y, float* xd, float* yd)
{   /* ApplyProject9 (M/MosaicWithoutPos.h:331-336): division form with h[8] */
    *xd = (h[0] * x + h[1] * y + h[2]) / (h[6] * x + h[7] * y + h[8]);
    *yd = (h[3] * x + h[4] * y + h[5]) / (h[6] * x + h[7] * y + h[8]);
}
int orc_paste(const float* H, int n, int img_w, int img_h, const uint8_t** srcs, int src_step,
              int* out_w, int* out_h, uint8_t* out, int out_step)
{
    float minX = (float)(1 << 29), minY = (float)(1 << 29), maxX = (float)(-(1 << 29)), maxY = (float)(-(1 << 29));
    float cx[4] = {0, (float)(img_w - 1), (float)(img_w - 1), 0}, cy[4] = {0, 0, (float)(img_h - 1), (float)(img_h - 1)};
    for (int k = 0; k < n; k++) {
        const float* m = H + (size_t)k * 9;
        if (m[8] == 0) continue;
        for (int i = 0; i < 4; i++) {
            float bx, by;
            apply_project9(m, cx[i], cy[i], &bx, &by);
            if (bx < minX) minX = bx;
            if (bx > maxX) maxX = bx;
            if (by < minY) minY = by;
            if (by > maxY) maxY = by;
        }
    }
    int W = (int)(maxX - minX + 1.5f), Hh = (int)(maxY - minY + 1.5f);
    *out_w = W; *out_h = Hh;
    if (!out) return 0;
    for (int y = 0; y < Hh; y++) memset(out + (size_t)y * out_step, 0, (size_t)W * 3);
    float dgx = -minX, dgy = -minY;
    int w1 = img_w - 1, h1 = img_h - 1;
    for (int k = 0; k < n; k++) {
        const float* m = H + (size_t)k * 9;
        if (m[8] == 0) continue;
        float inv[9]; memset(inv, 0, sizeof(inv));
        orc_inverse_matrix(m, 3, inv, 1e-12f);
        float bminx = (float)(1 << 29), bminy = (float)(1 << 29), bmaxx = (float)(-(1 << 29)), bmaxy = (float)(-(1 << 29));
        for (int i = 0; i < 4; i++) {
            float bx, by;
            apply_project9(m, cx[i], cy[i], &bx, &by);
            bx += (0 + dgx); by += (0 + dgy);
            if (bx < bminx) bminx = bx;
            if (bx > bmaxx) bmaxx = bx;
            if (by < bminy) bminy = by;
            if (by > bmaxy) bmaxy = by;
        }
        int begY = (int)(bminy - 0.5f), endY = (int)(bmaxy + 0.5f), begX = (int)(bminx - 0.5f), endX = (int)(bmaxx + 0.5f);
        const uint8_t* src = srcs[k];
        for (int yd = begY; yd <= endY; yd++) {
            if (yd < 0 || yd >= Hh) continue;                     /* the reference would write out of bounds */
            uint8_t* row = out + (size_t)yd * out_step;
            for (int xd = begX; xd <= endX; xd++) {
                if (xd < 0 || xd >= W) continue;
                float xm = xd - 0 - dgx, ym = yd - 0 - dgy;
                float xs, ys;
                apply_project9(inv, xm, ym, &xs, &ys);
                int ix = (int)xs, iy = (int)ys;
                float p = ys - iy, q = xs - ix;
                if ((ys < 0) || (ys >= h1)) continue;
                if ((xs < 0) || (xs >= w1)) continue;
                const uint8_t* t = src + (size_t)iy * src_step + 3 * ix;
                for (int c = 0; c < 3; c++) {
                    int g1 = t[c], g2 = t[c + 3], g3 = t[c + src_step], g4 = t[c + src_step + 3];
                    float v = g1 * (1 - p) * (1 - q) + g2 * (1 - p) * q + g3 * p * (1 - q) + g4 * p * q;
                    row[3 * xd + c] = (uint8_t)(int)v;
                }
            }
        }
    }
    return 0;
}

/* K7 (multi-band blend) lives in oracle_blend.c */
