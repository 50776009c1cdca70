/* Prelude for compiling the reference's numeric core with g++ (oracle/_ref build only).
 *
 * - removes MSVC __declspec
 * - makes the floating-point overloads of abs()/sqrt() visible before the reference
 *   headers are parsed (they call unqualified abs() on floats, M/matrix.h:183,191,203)
 * - forward-declares the templates the reference uses before declaring them
 *   (MSVC parses templates lazily; g++ does not)
 * - redirects srand/rand/time to a documented LCG so that the sample-index stream of
 *   Ransac2D (M/mosaicimage.h:1777,1801-1813) is reproducible and shared with the CUDA
 *   kernel: MSVC's rand(): s = s*214013 + 2531011; return (s>>16) & 0x7fff.
 */
#ifndef UAVM_REF_SHIM_PRELUDE_H
#define UAVM_REF_SHIM_PRELUDE_H
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <cstdio>
#include <ctime>
#include <vector>
#include <algorithm>
#include <iostream>
#include <stdint.h>
using namespace std;

#define __declspec(x)
#define _declspec(x)
#define veTemp vtTemp   /* typo in a never-instantiated template, M/Bitmap.h:229 */

/* thread_local: bench.py's CPU arm runs one Ransac2D per host thread, like the reference's worker threads */
extern thread_local uint32_t g_ref_seed;       /* seed to install at the next srand() */
extern thread_local uint32_t g_ref_state;
extern thread_local uint64_t g_ref_rand_calls; /* number of rand() calls since last srand() */
static inline void ref_srand_hook(unsigned) { g_ref_state = g_ref_seed; g_ref_rand_calls = 0; }
static inline int ref_rand_hook() {
    g_ref_state = g_ref_state * 214013u + 2531011u;
    g_ref_rand_calls++;
    return (int)((g_ref_state >> 16) & 0x7fff);
}
static inline long ref_time_hook() { return 0; }

template <class T> int TransposeMatrix(T* pSrc, int row, int col, T* pDst);
template <class T> int MulMatrix(T* pSrc1, int row1, int col1, T* pSrc2, int row2, int col2, T* pDst);
template <class T> int InverseMatrix(const T* pSrc, const int order, T* pDst, const T SMALL_NUMBER = 1e-6, int out = 0);
template <class T1, class T2, class T3> inline void ApplyProjectMat2(T1 xSrc, T1 ySrc, T2& xDst, T2& yDst, T3* M);
template <class T1, class T2, class T3> inline void ApplyAffineMat2(T1 xSrc, T1 ySrc, T2& xDst, T2& yDst, T3* M);
template <class T1, class T2, class T3> inline void DistanceOfTwoPoints(T1 x1, T1 y1, T2 x2, T2 y2, T3& dist);
#endif
