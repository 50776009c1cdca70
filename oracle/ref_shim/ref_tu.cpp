/* Translation unit that compiles the REFERENCE's own RANSAC numeric core (read in place from
 * /root/reference at build time; nothing is copied into this repository) into
 * oracle/_ref/libref_ransac.so.  TEST INFRASTRUCTURE ONLY.
 *
 * Sources pulled in:  M/matrix.h, M/LeastSquare.h, M/mvMath.h, M/Point.h, M/Bitmap.h
 * (+ M/MemoryPool.cpp compiled separately) and Ransac2D, extracted at build time by line
 * range (M/mosaicimage.h:24-34,1729-2035) into a temporary file that the Makefile deletes.
 */
#include "prelude.h"
#include "matrix.h"
#include "LeastSquare.h"
#include "mvMath.h"
#include "Bitmap.h"

/* redirect the RNG only for the Ransac2D body */
#define srand(x) ref_srand_hook(x)
#define rand() ref_rand_hook()
#define time(x) ref_time_hook()
#include "ransac2d_extract.inc"
#undef srand
#undef rand
#undef time

thread_local uint32_t g_ref_seed = 1;
thread_local uint32_t g_ref_state = 1;
thread_local uint64_t g_ref_rand_calls = 0;

extern "C" {

/* Ransac2D (M/mosaicimage.h:1729).  xy1/xy2: n x 2 floats.  Returns the function's bool.
 * inlier_mask[n] is derived from the ids of the returned inlier lists. */
int ref_ransac2d(const float* xy1, const float* xy2, int n, float ransac_dist, int sample_times,
                 uint32_t seed, unsigned char* inlier_mask, float H[9], int* n_inliers,
                 uint64_t* rand_calls)
{
    std::vector<pool::SfPoint> p1(n), p2(n), in1, in2;
    for (int i = 0; i < n; i++) {
        p1[i].x = xy1[2 * i]; p1[i].y = xy1[2 * i + 1]; p1[i].id = i;
        p2[i].x = xy2[2 * i]; p2[i].y = xy2[2 * i + 1]; p2[i].id = i;
    }
    for (int i = 0; i < 9; i++) H[i] = 0.f;
    g_ref_seed = seed;
    bool ok = Ransac2D(p1, p2, in1, in2, H, ransac_dist, sample_times);
    if (inlier_mask) {
        memset(inlier_mask, 0, n);
        for (size_t k = 0; k < in1.size(); k++) inlier_mask[in1[k].id] = 1;
    }
    if (n_inliers) *n_inliers = (int)in1.size();
    if (rand_calls) *rand_calls = g_ref_rand_calls;
    return ok ? 1 : 0;
}

/* SolveHomographyMatrix (M/matrix.h:783) */
int ref_solve_homography(const float* xy1, const float* xy2, int n, float h[9])
{
    std::vector<pool::SfPoint> p1(n), p2(n);
    for (int i = 0; i < n; i++) {
        p1[i].x = xy1[2 * i]; p1[i].y = xy1[2 * i + 1]; p1[i].id = i;
        p2[i].x = xy2[2 * i]; p2[i].y = xy2[2 * i + 1]; p2[i].id = i;
    }
    return SolveHomographyMatrix(&p1[0], &p2[0], n, h) ? 1 : 0;
}

/* NonlinearLeastSquareProjection2 (M/LeastSquare.h:353) */
int ref_nls_projection2(const float* xy1, const float* xy2, int n, float out[9], const float init[9],
                        float stop)
{
    std::vector<pool::SfPoint> p1(n), p2(n);
    for (int i = 0; i < n; i++) {
        p1[i].x = xy1[2 * i]; p1[i].y = xy1[2 * i + 1]; p1[i].id = i;
        p2[i].x = xy2[2 * i]; p2[i].y = xy2[2 * i + 1]; p2[i].id = i;
    }
    float init_copy[9];
    memcpy(init_copy, init, sizeof(init_copy));
    return NonlinearLeastSquareProjection2(&p1[0], &p2[0], n, out, init_copy, stop);
}

/* InverseMatrix (M/matrix.h:147) */
int ref_inverse_matrix(const float* src, int order, float* dst, float eps)
{
    return InverseMatrix(src, order, dst, eps, 0);
}
}
