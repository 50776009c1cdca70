// canvas.cu — device side of the warp stage of LaplacianPyramidBlending (M/MosaicImage.cpp:2205-2510):
//   K5  k5_warp_chips   bilinear inverse warp of every kept frame into its chip + validity mask (:2350-2448)
// (K6 seam masks live in masks.cu, K7 multi-band blend in blend.cu.)
//
// HBM layout: source frames are stored as BGRA (uchar4) so that one 32-bit load returns a whole pixel
// (a BGR triple straddles words; 12 byte loads per output pixel would make the kernel LSU-bound);
// chips are packed BGR u8 with a row step of 3*align4(chip_w) so every thread stores whole 12-byte
// groups; masks are u8 with a row step of align4(chip_w).
#include <string.h>
#include "canvas.h"

namespace {

// BGR (3 B/px, arbitrary step) -> BGRA (4 B/px)
__global__ void __launch_bounds__(256) k5_bgr_to_bgra(const uint8_t* __restrict__ src, int w, int h, int step,
                                                       uchar4* __restrict__ dst, int dst_step_px)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= w || y >= h) return;
    const uint8_t* s = src + (size_t)y * step + 3 * x;
    dst[(size_t)y * dst_step_px + x] = make_uchar4(s[0], s[1], s[2], 0);
}

// one channel of the reference's bilinear expression (:2398-2410), evaluated left to right in float:
//   uchar( g1*(1-p)*(1-q) + g2*(1-p)*q + g3*p*(1-q) + g4*p*q )
// u8 -> f32 conversion and the first multiply are ONE exact FFMA off the conversion (XU) pipe: the byte is
// placed in the mantissa of 2^23 (G = 2^23 + g, PRMT) and fma(G, w, -(2^23 * w)) rounds the infinitely
// precise value (2^23 + g) w - 2^23 w = g w once — bit-identical to float(g) * w (2^23 * w is exact).
// The remaining multiplies and adds stay separate instructions (-fmad=false), in the reference's order.
template <int K>
__device__ __forceinline__ uint32_t bilinear_channel(uint32_t t00, uint32_t t01, uint32_t t10, uint32_t t11,
                                                     float p, float q, float omp, float omq, float c_omp, float c_p,
                                                     uint32_t two23)
{
    const float G1 = __int_as_float(__byte_perm(t00, two23, 0x7540 | K));
    const float G2 = __int_as_float(__byte_perm(t01, two23, 0x7540 | K));
    const float G3 = __int_as_float(__byte_perm(t10, two23, 0x7540 | K));
    const float G4 = __int_as_float(__byte_perm(t11, two23, 0x7540 | K));
    const float a1 = __fmaf_rn(G1, omp, c_omp), a2 = __fmaf_rn(G2, omp, c_omp);      // g1*(1-p), g2*(1-p)
    const float a3 = __fmaf_rn(G3, p, c_p), a4 = __fmaf_rn(G4, p, c_p);              // g3*p, g4*p
    const float v = a1 * omq + a2 * q + a3 * omq + a4 * q;
    return (uint32_t)__float2int_rz(v);            // truncation, value in [0, 255]
}

constexpr int kWarpTileW = 128;   // 32 threads x 4 px
constexpr int kWarpRows = 4;      // rows per thread
constexpr int kWarpTileH = 8 * kWarpRows;

// AFFINE: inv[6] == inv[7] == 0 and inv[8] == 1, so the reference's denominator is exactly 1.0f for every
// pixel and x / 1.0f == x: the two divides per coordinate (:2359-2362) are skipped without changing a bit.
template <bool AFFINE>
__global__ void __launch_bounds__(256)
k5_warp_chips(const ChipDesc* __restrict__ descs, int img_w, int img_h, int src_step_px, float dgx, float dgy,
              uint32_t two23 /* = 0x4B000000, bits of 2^23: a kernel argument so PRMT takes the SELECTOR as its immediate */,
              float w1f, float h1f)
{
    const ChipDesc& D = descs[blockIdx.z];
    if (!D.keep || (D.affine != 0) != AFFINE) return;
    const int x0 = blockIdx.x * kWarpTileW + threadIdx.x * 4;
    const int ybase = blockIdx.y * kWarpTileH + threadIdx.y;
    if (x0 >= D.chip_w || blockIdx.y * kWarpTileH >= D.chip_h) return;
    const float iv0 = D.inv[0], iv1 = D.inv[1], iv2 = D.inv[2], iv3 = D.inv[3], iv4 = D.inv[4], iv5 = D.inv[5];
    const float iv6 = D.inv[6], iv7 = D.inv[7], iv8 = D.inv[8];
    const float w1 = w1f, h1 = h1f;                 // (float)(width-1), (float)(height-1): kernel arguments, not re-converted per pixel
    const uint32_t* __restrict__ src = reinterpret_cast<const uint32_t*>(D.src);
    const float fbx = (float)D.beg_x, fby = (float)D.beg_y, sx = D.sx, sy = D.sy;
    // xTemp = xDst - dGx - sx + begBoxX (:2356); (float)(x0 + i) == (float)x0 + i exactly (both < 2^24)
    float xt[4];
    const float fx0 = (float)x0;
#pragma unroll
    for (int i = 0; i < 4; i++) xt[i] = (fx0 + (float)i) - dgx - sx + fbx;
    uint8_t* chip_row = D.chip + (size_t)ybase * D.chip_step + 3 * x0;
    uint8_t* mask_row = D.mask + (size_t)ybase * D.mask_step + x0;
#pragma unroll
    for (int ry = 0; ry < kWarpRows; ry++) {
        const int yd = ybase + ry * 8;
        if (yd >= D.chip_h) break;
        const float yt = (float)yd - dgy - sy + fby;               // yTemp (:2357)
        const float ya = yt * iv1, yb = yt * iv4;
        uint32_t o[12];
        uint32_t mbits = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            float xs = xt[i] * iv0 + ya + iv2;
            float ys = xt[i] * iv3 + yb + iv5;
            if (!AFFINE) {
                const float den = xt[i] * iv6 + yt * iv7 + iv8;
                xs = xs / den;
                ys = ys / den;
            }
            o[3 * i] = 0; o[3 * i + 1] = 0; o[3 * i + 2] = 0;
            if ((xs >= 0.0f) && (xs < w1) && (ys >= 0.0f) && (ys < h1)) {
                const int iy = __float2int_rz(ys), ix = __float2int_rz(xs);
                const float p = ys - (float)iy, q = xs - (float)ix;
                const float omp = 1.0f - p, omq = 1.0f - q;
                const float c_omp = -8388608.0f * omp, c_p = -8388608.0f * p;        // exact (power-of-two scale)
                const uint32_t off = (uint32_t)(iy * src_step_px + ix);        // < 2^24 pixels per frame: 32-bit offset from the frame base
                const uint32_t t00 = __ldg(src + off), t01 = __ldg(src + off + 1u);
                const uint32_t t10 = __ldg(src + off + (uint32_t)src_step_px), t11 = __ldg(src + off + (uint32_t)src_step_px + 1u);
                o[3 * i] = bilinear_channel<0>(t00, t01, t10, t11, p, q, omp, omq, c_omp, c_p, two23);
                o[3 * i + 1] = bilinear_channel<1>(t00, t01, t10, t11, p, q, omp, omq, c_omp, c_p, two23);
                o[3 * i + 2] = bilinear_channel<2>(t00, t01, t10, t11, p, q, omp, omq, c_omp, c_p, two23);
                mbits |= 0xffu << (8 * i);
            }
        }
        uint32_t* crow = reinterpret_cast<uint32_t*>(chip_row + (size_t)ry * 8 * D.chip_step);
        // 12 bytes -> 3 words with byte permutes (3 PRMT per word instead of shift/or chains)
        crow[0] = __byte_perm(__byte_perm(o[0], o[1], 0x0040), __byte_perm(o[2], o[3], 0x0040), 0x5410);
        crow[1] = __byte_perm(__byte_perm(o[4], o[5], 0x0040), __byte_perm(o[6], o[7], 0x0040), 0x5410);
        crow[2] = __byte_perm(__byte_perm(o[8], o[9], 0x0040), __byte_perm(o[10], o[11], 0x0040), 0x5410);
        *reinterpret_cast<uint32_t*>(mask_row + (size_t)ry * 8 * D.mask_step) = mbits;
    }
}

inline int align4(int v) { return (v + 3) & ~3; }

}  // namespace

int uavm_canvas_upload_desc(uavm_ctx* ctx, uavm_canvas* cv)
{
    UAVM_CUDA(ctx, cudaMemcpyAsync(cv->d_desc, cv->desc.data(), cv->desc.size() * sizeof(ChipDesc), cudaMemcpyHostToDevice, ctx->stream));
    UAVM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return UAVM_OK;
}

extern "C" int uavm_canvas_create(uavm_ctx* ctx, int n_images, int img_w, int img_h, const float* H, const int32_t* keep, uavm_canvas** out)
{
    if (!ctx || !out || n_images <= 0 || img_w < 2 || img_h < 2 || !H) return UAVM_EINVAL;
    *out = nullptr;
    uavm_canvas* cv = new uavm_canvas();
    cv->n = n_images; cv->img_w = img_w; cv->img_h = img_h; cv->src_step_px = img_w;
    cv->H.assign(H, H + (size_t)n_images * 9);
    cv->chips.resize(n_images);
    int rc = uavm_canvas_layout_compute(H, keep, n_images, img_w, img_h, &cv->layout, cv->chips.data());
    if (rc != UAVM_OK) { delete cv; return rc; }
    if (cv->layout.canvas_w <= 0 || cv->layout.canvas_h <= 0 || (int64_t)cv->layout.canvas_w * cv->layout.canvas_h > (int64_t)1 << 33) {
        UAVM_SET_ERR(ctx, "canvas %d x %d out of range", cv->layout.canvas_w, cv->layout.canvas_h);
        delete cv; return UAVM_EINVAL;
    }
    cv->desc.resize(n_images);
    size_t chip_off = 0, mask_off = 0;
    std::vector<size_t> coff(n_images), moff(n_images);
    for (int k = 0; k < n_images; k++) {
        ChipDesc& d = cv->desc[k]; memset(&d, 0, sizeof(d));
        const uavm_chip_layout& c = cv->chips[k];
        d.keep = c.keep;
        if (!c.keep) continue;
        if (c.chip_w <= 0 || c.chip_h <= 0 || c.chip_w > (1 << 20) || c.chip_h > (1 << 20)) {
            UAVM_SET_ERR(ctx, "chip %d has size %d x %d", k, c.chip_w, c.chip_h);
            delete cv; return UAVM_EINVAL;
        }
        d.chip_w = c.chip_w; d.chip_h = c.chip_h;
        d.mask_step = align4(c.chip_w); d.chip_step = 3 * d.mask_step;
        d.beg_x = c.beg_x; d.beg_y = c.beg_y; d.sx = c.sx; d.sy = c.sy;
        memcpy(d.inv, c.inv, sizeof(d.inv)); memcpy(d.quad, c.quad, sizeof(d.quad));
        d.affine = (c.inv[6] == 0.0f && c.inv[7] == 0.0f && c.inv[8] == 1.0f) ? 1 : 0;
        coff[k] = chip_off; moff[k] = mask_off;
        chip_off += (size_t)d.chip_step * d.chip_h; mask_off += (size_t)d.mask_step * d.chip_h;
        chip_off = (chip_off + 255) & ~(size_t)255; mask_off = (mask_off + 255) & ~(size_t)255;
        if (c.chip_w > cv->max_chip_w) cv->max_chip_w = c.chip_w;
        if (c.chip_h > cv->max_chip_h) cv->max_chip_h = c.chip_h;
    }
    cv->chips_bytes = chip_off; cv->masks_bytes = mask_off;
    UAVM_CUDA(ctx, cudaSetDevice(ctx->device));
    UAVM_CUDA(ctx, cudaMalloc(&cv->d_src, (size_t)n_images * img_h * cv->src_step_px * sizeof(uchar4)));
    UAVM_CUDA(ctx, cudaMalloc(&cv->d_chips, chip_off + 256));
    UAVM_CUDA(ctx, cudaMalloc(&cv->d_masks, mask_off + 256));
    UAVM_CUDA(ctx, cudaMalloc(&cv->d_desc, (size_t)n_images * sizeof(ChipDesc)));
    cv->stage_bytes = (size_t)img_h * img_w * 3;
    UAVM_CUDA(ctx, cudaMalloc(&cv->d_stage, cv->stage_bytes));
    for (int k = 0; k < n_images; k++) {
        ChipDesc& d = cv->desc[k];
        d.src = cv->d_src + (size_t)k * img_h * cv->src_step_px;
        if (!d.keep) continue;
        d.chip = cv->d_chips + coff[k]; d.mask = cv->d_masks + moff[k];
    }
    cv->band_y0 = 0; cv->band_y1 = cv->layout.canvas_h; cv->band_Y0 = 0; cv->band_Y1 = (cv->layout.canvas_h + 31) & ~31;
    rc = uavm_canvas_upload_desc(ctx, cv);
    if (rc != UAVM_OK) { uavm_canvas_destroy(ctx, cv); return rc; }
    *out = cv;
    return UAVM_OK;
}

// Multi-GPU canvas sharding: this context computes only canvas rows [y0, y1) (+ halo rows on both sides that make
// the band interior bit-identical to the untiled blend, see blend.cu).  Chips that cannot touch the computed rows
// are deactivated: they are neither warped nor masked nor fed.  y0, y1, halo: multiples of 32 (y1 may be the
// canvas height).  y0 = 0, y1 = canvas_h, halo = 0 restores the whole canvas.
extern "C" int uavm_canvas_set_band(uavm_ctx* ctx, uavm_canvas* cv, int y0, int y1, int halo)
{
    if (!ctx || !cv) return UAVM_EINVAL;
    const int ch = cv->layout.canvas_h;
    if (y0 < 0 || y1 > ch || y0 >= y1 || halo < 0 || (y0 % 32) || (halo % 32) || ((y1 % 32) && y1 != ch)) {
        UAVM_SET_ERR(ctx, "set_band: rows must be multiples of 32 inside the canvas"); return UAVM_EINVAL;
    }
    const int Hpad = (ch + 31) & ~31;
    cv->band_y0 = y0; cv->band_y1 = y1;
    cv->band_Y0 = y0 - halo > 0 ? y0 - halo : 0;
    cv->band_Y1 = ((y1 + 31) & ~31) + halo < Hpad ? ((y1 + 31) & ~31) + halo : Hpad;
    cv->banded = !(y0 == 0 && y1 == ch);
    const int gap = 3 * 32 + 32;                        // feed ROI grows a chip by 3 * 2^5 rows, then aligns to 32
    cv->max_chip_w = 0; cv->max_chip_h = 0;
    for (int k = 0; k < cv->n; k++) {
        ChipDesc& d = cv->desc[k];
        const uavm_chip_layout& c = cv->chips[k];
        bool active = c.keep != 0;
        if (active && cv->banded) active = (c.beg_y - gap < cv->band_Y1) && (c.beg_y + c.chip_h + gap > cv->band_Y0);
        d.keep = active ? 1 : 0;
        if (active) { if (c.chip_w > cv->max_chip_w) cv->max_chip_w = c.chip_w; if (c.chip_h > cv->max_chip_h) cv->max_chip_h = c.chip_h; }
    }
    cv->nbr_dirty = true; cv->warped = false; cv->seamed = false; cv->blended = false;
    return uavm_canvas_upload_desc(ctx, cv);
}
extern "C" int uavm_canvas_is_active(uavm_canvas* cv, int image)
{
    if (!cv || image < 0 || image >= cv->n) return 0;
    return cv->desc[image].keep;
}

extern "C" void uavm_canvas_destroy(uavm_ctx* ctx, uavm_canvas* cv)
{
    if (!cv) return;
    if (ctx) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); }
    uavm_blend_free(cv);
    cudaFree(cv->d_src); cudaFree(cv->d_chips); cudaFree(cv->d_masks); cudaFree(cv->d_dist); cudaFree(cv->d_dist_max);
    cudaFree(cv->d_desc); cudaFree(cv->d_nbr); cudaFree(cv->d_stage); cudaFree(cv->d_result); cudaFree(cv->d_result_mask);
    delete cv;
}

extern "C" int uavm_canvas_get_layout(uavm_canvas* cv, uavm_canvas_layout* canvas, uavm_chip_layout* chips)
{
    if (!cv) return UAVM_EINVAL;
    if (canvas) *canvas = cv->layout;
    if (chips) memcpy(chips, cv->chips.data(), cv->chips.size() * sizeof(uavm_chip_layout));
    return UAVM_OK;
}

extern "C" int uavm_canvas_set_image(uavm_ctx* ctx, uavm_canvas* cv, int image, const uint8_t* bgr, int step, int is_device)
{
    if (!ctx || !cv || image < 0 || image >= cv->n || !bgr || step < cv->img_w * 3) return UAVM_EINVAL;
    const uint8_t* src = bgr; int sstep = step;
    if (!is_device) {
        UAVM_CUDA(ctx, cudaMemcpy2DAsync(cv->d_stage, (size_t)cv->img_w * 3, bgr, (size_t)step, (size_t)cv->img_w * 3, cv->img_h,
                                         cudaMemcpyHostToDevice, ctx->stream));
        src = cv->d_stage; sstep = cv->img_w * 3;
    }
    dim3 grid((cv->img_w + 255) / 256, cv->img_h);
    k5_bgr_to_bgra<<<grid, 256, 0, ctx->stream>>>(src, cv->img_w, cv->img_h, sstep,
                                                   cv->d_src + (size_t)image * cv->img_h * cv->src_step_px, cv->src_step_px);
    UAVM_CHECK_LAUNCH(ctx);
    return UAVM_OK;
}

extern "C" int uavm_canvas_warp(uavm_ctx* ctx, uavm_canvas* cv)
{
    if (!ctx || !cv) return UAVM_EINVAL;
    if (cv->max_chip_w <= 0) return UAVM_OK;
    dim3 grid((cv->max_chip_w + kWarpTileW - 1) / kWarpTileW, (cv->max_chip_h + kWarpTileH - 1) / kWarpTileH, cv->n);
    dim3 block(32, 8);
    bool any_affine = false, any_proj = false;
    for (int k = 0; k < cv->n; k++)
        if (cv->desc[k].keep) { if (cv->desc[k].affine) any_affine = true; else any_proj = true; }
    if (any_affine) {
        k5_warp_chips<true><<<grid, block, 0, ctx->stream>>>(cv->d_desc, cv->img_w, cv->img_h, cv->src_step_px, cv->layout.dgx, cv->layout.dgy, 0x4B000000u,
                                                             (float)(cv->img_w - 1), (float)(cv->img_h - 1));
        UAVM_CHECK_LAUNCH(ctx);
    }
    if (any_proj) {
        k5_warp_chips<false><<<grid, block, 0, ctx->stream>>>(cv->d_desc, cv->img_w, cv->img_h, cv->src_step_px, cv->layout.dgx, cv->layout.dgy, 0x4B000000u,
                                                             (float)(cv->img_w - 1), (float)(cv->img_h - 1));
        UAVM_CHECK_LAUNCH(ctx);
    }
    cv->warped = true;
    return UAVM_OK;
}

extern "C" int uavm_canvas_get_chip(uavm_ctx* ctx, uavm_canvas* cv, int image, uint8_t* chip_bgr, int chip_step, uint8_t* mask, int mask_step)
{
    if (!ctx || !cv || image < 0 || image >= cv->n) return UAVM_EINVAL;
    const ChipDesc& d = cv->desc[image];
    if (!d.keep) return UAVM_EINVAL;
    if (chip_bgr) {
        if (chip_step < d.chip_w * 3) return UAVM_EINVAL;
        UAVM_CUDA(ctx, cudaMemcpy2DAsync(chip_bgr, (size_t)chip_step, d.chip, (size_t)d.chip_step, (size_t)d.chip_w * 3, d.chip_h, cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (mask) {
        if (mask_step < d.chip_w) return UAVM_EINVAL;
        UAVM_CUDA(ctx, cudaMemcpy2DAsync(mask, (size_t)mask_step, d.mask, (size_t)d.mask_step, (size_t)d.chip_w, d.chip_h, cudaMemcpyDeviceToHost, ctx->stream));
    }
    UAVM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return UAVM_OK;
}
