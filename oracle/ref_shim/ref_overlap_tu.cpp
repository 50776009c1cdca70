/* Compiles the REFERENCE's own overlap filter (ResampleByOverlap and helpers, M/MosaicImage.cpp:1884-2201,
 * M/ImageMath.cpp:9-54,88-103,144-176,399-412, M/imageMath.h:26-87) in place from /root/reference into
 * oracle/_ref/libref_ransac.so.  TEST INFRASTRUCTURE ONLY; the extracts are temporary files deleted by the
 * Makefile. */
#include "prelude.h"
#include "Point.h"
using namespace pool;
namespace pool { const float pi = 3.1415926f; struct Distance { float dist; int seq; bool operator<(const Distance& r) const { return dist < r.dist; } }; }
struct _IplImage { int width, height; };
typedef _IplImage IplImage;
struct ProjectMat { float m[9]; };
template<class T1, class T2, class T3> inline void DistanceOfTwoPoints(T1 x1, T1 y1, T2 x2, T2 y2, T3& dist)
{ dist = sqrt((T3)(x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2)); }
#include "imagemath_h_extract.inc"
#include "imagemath_cpp_extract.inc"
#include "overlap_extract.inc"

extern "C" int ref_resample_by_overlap(const float* H, int n, int w, int h, float overlap_t, int* keep)
{
    std::vector<IplImage> imgs(n); std::vector<IplImage*> ptrs(n);
    std::vector<ProjectMat> T(n);
    for (int i = 0; i < n; i++) { imgs[i].width = w; imgs[i].height = h; ptrs[i] = &imgs[i]; memcpy(T[i].m, H + 9 * i, 36); }
    std::vector<int> ab;
    ResampleByOverlap(&ptrs[0], n, overlap_t, &T[0], ab);
    for (int i = 0; i < n; i++) keep[i] = ab[i];
    return 0;
}
