// canvas_host.cpp — host-side geometry of the warp/blend stage: canvas sizing and per-image chip boxes
// (LaplacianPyramidBlending, M/MosaicImage.cpp:2233-2348) and the overlap filter (ResampleByOverlap,
// M/MosaicImage.cpp:2070-2201 with its helpers :1884-2067 and M/ImageMath.cpp:9-54,88-103,144-176,399-412).
// These are tiny sequential float computations (O(N) / O(N^2) over images) and stay on the host, in the
// reference's evaluation order; compiled by nvcc's host compiler with contraction off (-ffp-contract is
// irrelevant on x86-64 without -mfma; build uses plain SSE2 scalar float).
#include <math.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include "internal.h"

namespace {

// InverseMatrix (M/matrix.h:147-296) for small orders; same quirks as the device version.
int inverse_small(const float* src, int order, float* dst, float eps)
{
    if (order > 13 || order < 2) return -1;
    float T[400];
    bool used[16];
    const int o2 = order * 2;
    for (int i = 0; i < order * o2; i++) T[i] = 0.0f;
    for (int i = 0; i < order; i++) {
        used[i] = false;
        T[i * o2 + order + i] = 1.0f;
        for (int j = 0; j < order; j++) T[i * o2 + j] = src[i * order + j];
    }
    for (int i = 0; i < order; i++) {
        float e = 0.0f; int row = 0;
        for (int j = 0; j < order; j++) {
            if (used[j]) continue;
            if (fabsf(T[j * o2 + i]) > eps) { used[j] = true; e = T[j * o2 + i]; row = j; break; }
        }
        if (fabsf(e) < eps) return 0;
        for (int c = 0; c < o2; c++) T[row * o2 + c] = T[row * o2 + c] / e;
        for (int j = 0; j < order; j++) {
            if (j == row) continue;
            if (fabsf(T[j * o2 + i]) < eps) continue;
            const float nf = -T[j * o2 + i];
            for (int c = 0; c < o2; c++) T[j * o2 + c] = T[j * o2 + c] + nf * T[row * o2 + c];
        }
    }
    for (int r = 0; r < order; r++) {
        int target = -1;
        for (int i = 0; i < order && target < 0; i++)
            if (T[i * o2 + r] == 1.0f) target = i;
        if (target >= 0 && target != r)
            for (int j = 0; j < o2; j++) std::swap(T[r * o2 + j], T[target * o2 + j]);
    }
    for (int i = 0; i < order; i++)
        for (int j = 0; j < order; j++) dst[i * order + j] = T[i * o2 + order + j];
    return 1;
}

}  // namespace

// canvas bbox over kept quads (min/max start at 0, so the canvas always contains the reference origin,
// :2233), newW = int(maxX - minX + 1.5f) (:2292), integer chip box [int(beg), int(end + 0.5f)] and
// sub-pixel shift s = int(beg) - beg (:2314-2325), quad corners in chip coordinates (:2327-2337),
// InverseMatrix(H, 3, inv, 1e-12f) (:2347-2348).
extern "C" int uavm_canvas_layout_compute(const float* H, const int32_t* keep, int n, int img_w, int img_h,
                                          uavm_canvas_layout* canvas, uavm_chip_layout* chips)
{
    if (!H || n <= 0 || img_w < 2 || img_h < 2 || !canvas || !chips) return UAVM_EINVAL;
    float maxX = 0, maxY = 0, minX = 0, minY = 0;
    std::vector<float> beg(2 * (size_t)n, 0.0f), end(2 * (size_t)n, 0.0f);
    const float cx[4] = {0.0f, (float)(img_w - 1), (float)(img_w - 1), 0.0f};
    const float cy[4] = {0.0f, 0.0f, (float)(img_h - 1), (float)(img_h - 1)};
    for (int k = 0; k < n; k++) {
        const float* m = H + (size_t)k * 9;
        memset(&chips[k], 0, sizeof(chips[k]));
        if (keep && keep[k] == 0) continue;
        if (m[8] == 0) continue;                               // "skip me" sentinel (:2243)
        chips[k].keep = 1;
        float bmaxx = -536870912.0f, bmaxy = -536870912.0f, bminx = 536870912.0f, bminy = 536870912.0f;
        for (int i = 0; i < 4; i++) {
            const float xs = cx[i], ys = cy[i];
            const float xd = (xs * m[0] + ys * m[1] + m[2]) / (xs * m[6] + ys * m[7] + m[8]);
            const float yd = (xs * m[3] + ys * m[4] + m[5]) / (xs * m[6] + ys * m[7] + m[8]);
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
    const float dgx = -minX, dgy = -minY;
    canvas->dgx = dgx; canvas->dgy = dgy;
    canvas->canvas_w = (int)(maxX - minX + 1.5f);
    canvas->canvas_h = (int)(maxY - minY + 1.5f);
    for (int k = 0; k < n; k++) {
        if (!chips[k].keep) continue;
        const float* m = H + (size_t)k * 9;
        const float bx = beg[2 * k] + dgx, by = beg[2 * k + 1] + dgy;
        const float ex = end[2 * k] + dgx, ey = end[2 * k + 1] + dgy;
        const int ibx = (int)bx, iby = (int)by, iex = (int)(ex + 0.5f), iey = (int)(ey + 0.5f);
        const float sx = ibx - bx, sy = iby - by;
        chips[k].beg_x = ibx; chips[k].beg_y = iby;
        chips[k].chip_w = iex - ibx + 1; chips[k].chip_h = iey - iby + 1;
        chips[k].sx = sx; chips[k].sy = sy;
        for (int i = 0; i < 4; i++) {
            const float inv = 1 / (m[6] * cx[i] + m[7] * cy[i] + m[8]);          // ApplyProjectMat9 (M/matrix.h:1015-1024)
            const float tx = (m[0] * cx[i] + m[1] * cy[i] + m[2]) * inv;
            const float ty = (m[3] * cx[i] + m[4] * cy[i] + m[5]) * inv;
            chips[k].quad[2 * i] = tx + dgx + sx - ibx;
            chips[k].quad[2 * i + 1] = ty + dgy + sy - iby;
        }
        inverse_small(m, 3, chips[k].inv, 1e-12f);             // on failure inv stays zero
    }
    return UAVM_OK;
}
