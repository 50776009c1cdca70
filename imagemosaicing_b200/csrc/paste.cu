// paste.cu — the blending != 2 variant of the last stage: MosaicImagesRefined
// (M/MosaicWithoutPos.cpp:2194-2352).  No chips, no masks: every valid image is resampled straight into the
// mosaic in index order, so the LAST image covering a pixel wins.  Its canvas box differs from the blend
// path's (min/max start at +-2^29, no sub-pixel chip shift), so the geometry is computed here.
#include <math.h>
#include <string.h>
#include "canvas.h"

namespace {

struct PasteDesc { float inv[9]; int beg_x, beg_y, end_x, end_y; };

template <bool BGR>       // BGR: packed rows of src_step_px BYTES (canvas.h: src_bgr), else BGRA pixels
__global__ void __launch_bounds__(256)
k5_paste_image(const uchar4* __restrict__ src, int img_w, int img_h, int src_step_px, PasteDesc D, float dgx, float dgy,
               uint8_t* __restrict__ out, uint8_t* __restrict__ out_mask, int W, int H)
{
    const int xd = D.beg_x + blockIdx.x * blockDim.x + threadIdx.x;
    const int yd = D.beg_y + blockIdx.y;
    if (xd > D.end_x || yd > D.end_y || xd < 0 || yd < 0 || xd >= W || yd >= H) return;
    const float xm = (float)(xd - 0) - dgx, ym = (float)(yd - 0) - dgy;                      // xMid, yMid (:2299-2300)
    // ApplyProject9 (M/MosaicWithoutPos.h:331-336): two true divides per coordinate
    const float xs = (D.inv[0] * xm + D.inv[1] * ym + D.inv[2]) / (D.inv[6] * xm + D.inv[7] * ym + D.inv[8]);
    const float ys = (D.inv[3] * xm + D.inv[4] * ym + D.inv[5]) / (D.inv[6] * xm + D.inv[7] * ym + D.inv[8]);
    if ((ys < 0.0f) || (ys >= (float)(img_h - 1))) return;
    if ((xs < 0.0f) || (xs >= (float)(img_w - 1))) return;
    const int ix = __float2int_rz(xs), iy = __float2int_rz(ys);
    const float p = ys - (float)iy, q = xs - (float)ix;
    const float omp = 1.0f - p, omq = 1.0f - q;
    uchar4 t00, t01, t10, t11;
    if (BGR) {
        const uint8_t* r8 = reinterpret_cast<const uint8_t*>(src) + (size_t)iy * src_step_px + 3 * (size_t)ix;
        t00 = make_uchar4(r8[0], r8[1], r8[2], 0); t01 = make_uchar4(r8[3], r8[4], r8[5], 0);
        r8 += src_step_px;
        t10 = make_uchar4(r8[0], r8[1], r8[2], 0); t11 = make_uchar4(r8[3], r8[4], r8[5], 0);
    } else {
        const uchar4* r0 = src + (size_t)iy * src_step_px + ix;
        t00 = r0[0]; t01 = r0[1]; t10 = r0[src_step_px]; t11 = r0[src_step_px + 1];
    }
    uint8_t* d = out + ((size_t)yd * W + xd) * 3;
    d[0] = (uint8_t)__float2int_rz((float)t00.x * omp * omq + (float)t01.x * omp * q + (float)t10.x * p * omq + (float)t11.x * p * q);
    d[1] = (uint8_t)__float2int_rz((float)t00.y * omp * omq + (float)t01.y * omp * q + (float)t10.y * p * omq + (float)t11.y * p * q);
    d[2] = (uint8_t)__float2int_rz((float)t00.z * omp * omq + (float)t01.z * omp * q + (float)t10.z * p * omq + (float)t11.z * p * q);
    out_mask[(size_t)yd * W + xd] = 255;
}

int inverse3(const float* src, float* dst, float eps)      // InverseMatrix (M/matrix.h:147-296), order 3
{
    float T[3][6]; bool used[3] = {false, false, false};
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { T[i][j] = src[i * 3 + j]; T[i][3 + j] = (i == j) ? 1.0f : 0.0f; }
    for (int i = 0; i < 3; i++) {
        float e = 0.0f; int row = 0;
        for (int j = 0; j < 3; j++) { if (used[j]) continue; if (fabsf(T[j][i]) > eps) { used[j] = true; e = T[j][i]; row = j; break; } }
        if (fabsf(e) < eps) return 0;
        for (int c = 0; c < 6; c++) T[row][c] = T[row][c] / e;
        for (int j = 0; j < 3; j++) {
            if (j == row || fabsf(T[j][i]) < eps) continue;
            const float nf = -T[j][i];
            for (int c = 0; c < 6; c++) T[j][c] = T[j][c] + nf * T[row][c];
        }
    }
    for (int r = 0; r < 3; r++) {
        int target = -1;
        for (int i = 0; i < 3 && target < 0; i++) if (T[i][r] == 1.0f) target = i;
        if (target >= 0 && target != r) for (int c = 0; c < 6; c++) { float t = T[r][c]; T[r][c] = T[target][c]; T[target][c] = t; }
    }
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) dst[i * 3 + j] = T[i][3 + j];
    return 1;
}

void project9(const float* h, float x, float y, float& xd, float& yd)
{
    xd = (h[0] * x + h[1] * y + h[2]) / (h[6] * x + h[7] * y + h[8]);
    yd = (h[3] * x + h[4] * y + h[5]) / (h[6] * x + h[7] * y + h[8]);
}

}  // namespace

extern "C" int uavm_canvas_paste(uavm_ctx* ctx, uavm_canvas* cv)
{
    if (!ctx || !cv) return UAVM_EINVAL;
    UAVM_CUDA(ctx, cudaSetDevice(ctx->device));
    const int n = cv->n, w = cv->img_w, h = cv->img_h;
    const float cx[4] = {0.0f, (float)(w - 1), (float)(w - 1), 0.0f}, cy[4] = {0.0f, 0.0f, (float)(h - 1), (float)(h - 1)};
    float minX = 536870912.0f, minY = 536870912.0f, maxX = -536870912.0f, maxY = -536870912.0f;
    int n_valid = 0;
    for (int k = 0; k < n; k++) {
        const float* m = &cv->H[(size_t)k * 9];
        if (m[8] == 0) continue;
        n_valid++;
        for (int i = 0; i < 4; i++) {
            float bx, by;
            project9(m, cx[i], cy[i], bx, by);
            if (bx < minX) minX = bx;
            if (bx > maxX) maxX = bx;
            if (by < minY) minY = by;
            if (by > maxY) maxY = by;
        }
    }
    if (n_valid == 0) { UAVM_SET_ERR(ctx, "paste: no valid image"); return UAVM_EFAIL; }
    const int W = (int)(maxX - minX + 1.5f), H = (int)(maxY - minY + 1.5f);
    if (W <= 0 || H <= 0 || (int64_t)W * H > ((int64_t)1 << 33)) { UAVM_SET_ERR(ctx, "paste: canvas %d x %d out of range", W, H); return UAVM_EINVAL; }
    if (cv->result_w != W || cv->result_h != H) {
        cudaFree(cv->d_result); cudaFree(cv->d_result_mask); cv->d_result = nullptr; cv->d_result_mask = nullptr;
        UAVM_CUDA(ctx, cudaMalloc(&cv->d_result, (size_t)W * H * 3));
        UAVM_CUDA(ctx, cudaMalloc(&cv->d_result_mask, (size_t)W * H));
        cv->result_w = W; cv->result_h = H;
    }
    UAVM_CUDA(ctx, cudaMemsetAsync(cv->d_result, 0, (size_t)W * H * 3, ctx->stream));
    UAVM_CUDA(ctx, cudaMemsetAsync(cv->d_result_mask, 0, (size_t)W * H, ctx->stream));
    const float dgx = -minX, dgy = -minY;
    for (int k = 0; k < n; k++) {                                  // index order: the last image wins
        const float* m = &cv->H[(size_t)k * 9];
        if (m[8] == 0) continue;
        PasteDesc D; memset(&D, 0, sizeof(D));
        inverse3(m, D.inv, 1e-12f);
        float bminx = 536870912.0f, bminy = 536870912.0f, bmaxx = -536870912.0f, bmaxy = -536870912.0f;
        for (int i = 0; i < 4; i++) {
            float bx, by;
            project9(m, cx[i], cy[i], bx, by);
            bx += (0 + dgx); by += (0 + dgy);
            if (bx < bminx) bminx = bx;
            if (bx > bmaxx) bmaxx = bx;
            if (by < bminy) bminy = by;
            if (by > bmaxy) bmaxy = by;
        }
        D.beg_y = (int)(bminy - 0.5f); D.end_y = (int)(bmaxy + 0.5f);
        D.beg_x = (int)(bminx - 0.5f); D.end_x = (int)(bmaxx + 0.5f);
        if (D.end_x < D.beg_x || D.end_y < D.beg_y) continue;
        dim3 grid((D.end_x - D.beg_x + 256) / 256, D.end_y - D.beg_y + 1);
        if (cv->src_bgr)
            k5_paste_image<true><<<grid, 256, 0, ctx->stream>>>(reinterpret_cast<const uchar4*>(reinterpret_cast<const uint8_t*>(cv->d_src) + (size_t)k * h * cv->src_step_px),
                                                                w, h, cv->src_step_px, D, dgx, dgy, cv->d_result, cv->d_result_mask, W, H);
        else
            k5_paste_image<false><<<grid, 256, 0, ctx->stream>>>(cv->d_src + (size_t)k * h * cv->src_step_px, w, h, cv->src_step_px, D, dgx, dgy,
                                                                 cv->d_result, cv->d_result_mask, W, H);
        UAVM_CHECK_LAUNCH(ctx);
    }
    cv->blended = true;
    return UAVM_OK;
}
