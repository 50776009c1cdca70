/* Compiles the REFERENCE's own chip-warp loop (M/MosaicImage.cpp:2350-2448, the body of
 * LaplacianPyramidBlending's per-image loop) and FindMasksByDistMap (M/MosaicImage.cpp:1761-1881) in place
 * from /root/reference into oracle/_ref/libref_ransac.so.  TEST INFRASTRUCTURE ONLY.
 * The OpenCV types they touch are replaced by a minimal IplImage POD (the five fields read) and cvZero;
 * the one MSVC-only construct in the loop, the functional cast `unsigned char(expr)`, is rewritten by sed
 * to `(unsigned char)(expr)` at extract time (see oracle/Makefile).  Extracts are temporary files. */
#include "prelude.h"
#include "Point.h"
using namespace pool;
struct _IplImage { int nChannels; int width; int height; int widthStep; char* imageData; };
typedef _IplImage IplImage;
namespace cv { struct Point { int x, y; }; }
struct Rectangle4Points { SfPoint pt[4]; };
#define _IN
static inline void cvZero(IplImage* p) { memset(p->imageData, 0, (size_t)p->widthStep * p->height); }
void LineOf2Points1(float& a, float& b, float& c, float x1, float y1, float x2, float y2);   /* M/ImageMath.cpp:88 (compiled in ref_overlap_tu.cpp) */
#include "masks_extract.inc"

extern "C" void ref_warp_chip(const unsigned char* src, int w, int h, int step, float dGx, float dGy, float sx, float sy,
                              int begBoxX, int begBoxY, const float* inv, int wChip, int hChip,
                              unsigned char* chip, int chip_step, unsigned char* mask, int mask_step)
{
    IplImage img = {3, w, h, step, (char*)src};
    IplImage* arr[1] = {&img};
    IplImage** pImages = arr;
    const int n = 0;
    IplImage chipI = {3, wChip, hChip, chip_step, (char*)chip}, *pChipImage = &chipI;
    IplImage maskI = {1, wChip, hChip, mask_step, (char*)mask}, *pMask = &maskI;
    memset(pMask->imageData, 255, pMask->widthStep * pMask->height);          /* :2342 */
    int wsNew = pChipImage->widthStep;
    float pInvM[9];
    memcpy(pInvM, inv, sizeof(pInvM));
#include "warp_loop_extract.inc"
}

extern "C" int ref_find_masks(unsigned char** masks, const int* mask_step, const int* chip_w, const int* chip_h,
                              const float* quads /* n x 8 */, const int* tl_xy /* n x 2 */, int n, int rectW, int rectH)
{
    std::vector<IplImage> im(n); std::vector<IplImage*> ptr(n);
    std::vector<Rectangle4Points> rc(n); std::vector<cv::Point> tl(n);
    for (int i = 0; i < n; i++) {
        im[i].nChannels = 1; im[i].width = chip_w[i]; im[i].height = chip_h[i]; im[i].widthStep = mask_step[i]; im[i].imageData = (char*)masks[i];
        ptr[i] = &im[i];
        for (int k = 0; k < 4; k++) { rc[i].pt[k].x = quads[8 * i + 2 * k]; rc[i].pt[k].y = quads[8 * i + 2 * k + 1]; }
        tl[i].x = tl_xy[2 * i]; tl[i].y = tl_xy[2 * i + 1];
    }
    return FindMasksByDistMap(&ptr[0], n, &rc[0], &tl[0], rectW, rectH);
}
