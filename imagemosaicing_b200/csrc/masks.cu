// masks.cu — K6: distance-map seam masks.  Replaces FindMasksByDistMap (M/MosaicImage.cpp:1761-1881).
//
// Reference: per image, d(x,y) = min over the 4 quad edges of |A x + B y + C| / sqrt(A^2+B^2) for valid
// chip pixels, divided by the image's maximum; then every canvas pixel is owned by the image with the
// largest normalised distance (strict >, so the lowest image index wins ties and 0 never wins) and only
// the owner's mask is 255.
// Here: k6_dist_map (per chip pixel, + per-image max by atomicMax on the float bits), k6_normalize
// (true division by the max, in place), k6_owner (per chip pixel: am I the first arg-max among the chips
// whose boxes intersect mine?).  The O(N) scan over all images per canvas pixel becomes a scan over the
// chip's box-intersection neighbours, with the same first-wins rule.
#include <math.h>
#include <string.h>
#include "canvas.h"

namespace {

constexpr int kTileW = 128, kTileH = 8;

// Validity is recomputed from the pixel coordinates with exactly K5's expression chain (M/MosaicImage.cpp:2356-2362, :2370)
// instead of being read back from the chip: the pass is compute-only for rows this context does not hold (a band-sharded
// canvas warps and masks only its rows, yet the per-image maximum needs every pixel of the chip), and it saves the read.
// Rows outside canvas rows [row0, row1) are not stored.
__global__ void __launch_bounds__(256)
k6_dist_map(const ChipDesc* __restrict__ descs, float* __restrict__ dist_max, float dgx, float dgy, float w1f, float h1f, int row0, int row1)
{
    const ChipDesc& D = descs[blockIdx.z];
    if (!D.keep) return;
    const int x0 = blockIdx.x * kTileW + threadIdx.x * 4;
    const int r = blockIdx.y * kTileH + threadIdx.y;
    if (blockIdx.x * kTileW >= D.chip_w || blockIdx.y * kTileH >= D.chip_h) return;
    float mx = 0.0f;
    if (x0 < D.chip_w && r < D.chip_h) {
        const float yt = (float)r - dgy - D.sy + (float)D.beg_y;          // yTemp (:2357)
        const float ya = yt * D.inv[1], yb = yt * D.inv[4];
        float out[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int c = x0 + i;
            const float xt = (float)c - dgx - D.sx + (float)D.beg_x;      // xTemp (:2356)
            float xs = xt * D.inv[0] + ya + D.inv[2];
            float ys = xt * D.inv[3] + yb + D.inv[5];
            if (!D.affine) {
                const float den = xt * D.inv[6] + yt * D.inv[7] + D.inv[8];
                xs = xs / den; ys = ys / den;
            }
            const bool valid = (xs >= 0.0f) && (xs < w1f) && (ys >= 0.0f) && (ys < h1f);
            float mind = 0.0f;
            if (valid && c < D.chip_w) {
                mind = 536870912.0f;                                   // float minDist = 1<<29
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const float d = fabsf(D.lineA[e] * (float)c + D.lineB[e] * (float)r + D.lineC[e]) * D.lineInv[e];
                    if (d < mind) mind = d;
                }
                if (mind > mx) mx = mind;
            }
            out[i] = mind;
        }
        const int gy = r + D.beg_y;
        if (gy >= row0 && gy < row1)
            *reinterpret_cast<float4*>(D.dist + (size_t)r * D.mask_step + x0) = make_float4(out[0], out[1], out[2], out[3]);
    }
    // block max -> one atomic per block (distances are >= 0, so the uint order of the bits is the float order)
    __shared__ float smax[8];
    mx = __int_as_float(__reduce_max_sync(0xffffffffu, __float_as_int(mx)));
    if (threadIdx.x == 0) smax[threadIdx.y] = mx;
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        float m = smax[0];
#pragma unroll
        for (int i = 1; i < 8; i++) m = fmaxf(m, smax[i]);
        if (m > 0.0f) atomicMax(reinterpret_cast<unsigned int*>(dist_max + blockIdx.z), __float_as_uint(m));
    }
}

__global__ void __launch_bounds__(256)
k6_normalize(const ChipDesc* __restrict__ descs, const float* __restrict__ dist_max, int row0, int row1)
{
    const ChipDesc& D = descs[blockIdx.z];
    if (!D.keep) return;
    const int x0 = blockIdx.x * kTileW + threadIdx.x * 4;
    const int r = blockIdx.y * kTileH + threadIdx.y;
    if (x0 >= D.chip_w || r >= D.chip_h || r + D.beg_y < row0 || r + D.beg_y >= row1) return;
    const float mx = dist_max[blockIdx.z];
    float4* p = reinterpret_cast<float4*>(D.dist + (size_t)r * D.mask_step + x0);
    float4 v = *p;
    v.x = v.x / mx; v.y = v.y / mx; v.z = v.z / mx; v.w = v.w / mx;     // pMapRow[c] /= maxDist (:1830)
    *p = v;
}

__global__ void __launch_bounds__(256)
k6_owner(const ChipDesc* __restrict__ descs, const int32_t* __restrict__ nbr, int row0, int row1)
{
    const int n = blockIdx.z;
    const ChipDesc& D = descs[n];
    if (!D.keep) return;
    const int x0 = blockIdx.x * kTileW + threadIdx.x * 4;
    const int r = blockIdx.y * kTileH + threadIdx.y;
    if (x0 >= D.chip_w || r >= D.chip_h || r + D.beg_y < row0 || r + D.beg_y >= row1) return;
    const float4 own4 = *reinterpret_cast<const float4*>(D.dist + (size_t)r * D.mask_step + x0);
    const float own[4] = {own4.x, own4.y, own4.z, own4.w};
    bool win[4];
#pragma unroll
    for (int i = 0; i < 4; i++) win[i] = (own[i] > 0.0f) && (x0 + i < D.chip_w);
    const int gy = r + D.beg_y;
    for (int k = 0; k < D.nbr_cnt; k++) {
        const int m = nbr[D.nbr_off + k];
        const ChipDesc& E = descs[m];
        const int yc = gy - E.beg_y;
        if (yc < 0 || yc >= E.chip_h) continue;
        const float* erow = E.dist + (size_t)yc * E.mask_step;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int xc = x0 + i + D.beg_x - E.beg_x;
            if (!win[i] || xc < 0 || xc >= E.chip_w) continue;
            const float cur = erow[xc];
            if (m < n) { if (cur >= own[i]) win[i] = false; }        // an earlier image already holds the maximum
            else       { if (cur > own[i]) win[i] = false; }         // a later image strictly exceeds it
        }
    }
    const uint32_t bits = (win[0] ? 0xffu : 0u) | (win[1] ? 0xff00u : 0u) | (win[2] ? 0xff0000u : 0u) | (win[3] ? 0xff000000u : 0u);
    *reinterpret_cast<uint32_t*>(D.mask + (size_t)r * D.mask_step + x0) = bits;
}

// LineOf2Points1 (M/ImageMath.cpp:88-103)
void line_of_2_points(float& a, float& b, float& c, float x1, float y1, float x2, float y2)
{
    if (fabs(x1 - x2) < 0.000001) { a = 1.0f; b = 0; c = -x1; }
    else { a = (y1 - y2) / (x1 - x2); b = -1.0f; c = y1 - a * x1; }
}

}  // namespace

extern "C" int uavm_canvas_seam_masks(uavm_ctx* ctx, uavm_canvas* cv)
{
    if (!ctx || !cv) return UAVM_EINVAL;
    if (!cv->warped) { UAVM_SET_ERR(ctx, "seam_masks before warp"); return UAVM_EINVAL; }
    if (cv->max_chip_w <= 0) return UAVM_OK;
    UAVM_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!cv->d_dist) {
        UAVM_CUDA(ctx, cudaMalloc(&cv->d_dist, cv->masks_bytes * sizeof(float) + 1024));
        UAVM_CUDA(ctx, cudaMalloc(&cv->d_dist_max, (size_t)cv->n * sizeof(float)));
        cv->nbr_dirty = true;
    }
    if (cv->nbr_dirty) {
        cudaFree(cv->d_nbr); cv->d_nbr = nullptr;
        // edge lines, distance-map pointers and box-intersection neighbour lists
        std::vector<int32_t> nbr;
        for (int k = 0; k < cv->n; k++) {
            ChipDesc& d = cv->desc[k];
            if (!d.keep) continue;
            d.dist = cv->d_dist + (size_t)(d.mask - cv->d_masks);
            for (int e = 0; e < 4; e++) {
                const int f = (e + 1) & 3;
                line_of_2_points(d.lineA[e], d.lineB[e], d.lineC[e], d.quad[2 * e], d.quad[2 * e + 1], d.quad[2 * f], d.quad[2 * f + 1]);
                d.lineInv[e] = 1 / sqrtf(d.lineA[e] * d.lineA[e] + d.lineB[e] * d.lineB[e]);
            }
            d.nbr_off = (int32_t)nbr.size();
            for (int m = 0; m < cv->n; m++) {
                const ChipDesc& e = cv->desc[m];
                if (m == k || !e.keep) continue;
                if (e.beg_x < d.beg_x + d.chip_w && d.beg_x < e.beg_x + e.chip_w &&
                    e.beg_y < d.beg_y + d.chip_h && d.beg_y < e.beg_y + e.chip_h) nbr.push_back(m);
            }
            d.nbr_cnt = (int32_t)nbr.size() - d.nbr_off;
        }
        UAVM_CUDA(ctx, cudaMalloc(&cv->d_nbr, (nbr.size() + 1) * sizeof(int32_t)));
        if (!nbr.empty())
            UAVM_CUDA(ctx, cudaMemcpyAsync(cv->d_nbr, nbr.data(), nbr.size() * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
        int rc = uavm_canvas_upload_desc(ctx, cv);
        if (rc != UAVM_OK) return rc;
        cv->nbr_dirty = false;
    }
    UAVM_CUDA(ctx, cudaMemsetAsync(cv->d_dist_max, 0, (size_t)cv->n * sizeof(float), ctx->stream));
    dim3 grid((cv->max_chip_w + kTileW - 1) / kTileW, (cv->max_chip_h + kTileH - 1) / kTileH, cv->n);
    dim3 block(32, 8);
    // canvas rows this context holds (band + halo; the whole canvas when it is not sharded): the only rows K7 feeds from
    const int row0 = cv->banded ? cv->band_Y0 : 0, row1 = cv->banded ? cv->band_Y1 : cv->layout.canvas_h + 64;
    k6_dist_map<<<grid, block, 0, ctx->stream>>>(cv->d_desc, cv->d_dist_max, cv->layout.dgx, cv->layout.dgy, (float)(cv->img_w - 1), (float)(cv->img_h - 1), row0, row1);
    UAVM_CHECK_LAUNCH(ctx);
    k6_normalize<<<grid, block, 0, ctx->stream>>>(cv->d_desc, cv->d_dist_max, row0, row1);
    UAVM_CHECK_LAUNCH(ctx);
    k6_owner<<<grid, block, 0, ctx->stream>>>(cv->d_desc, cv->d_nbr, row0, row1);
    UAVM_CHECK_LAUNCH(ctx);
    cv->seamed = true; cv->mask_plane_valid = true;
    return UAVM_OK;
}
