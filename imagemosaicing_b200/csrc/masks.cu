// masks.cu — K6: distance-map seam masks.  Replaces FindMasksByDistMap (M/MosaicImage.cpp:1761-1881).
//
// Reference: per image, d(x,y) = min over the 4 quad edges of |A x + B y + C| / sqrt(A^2+B^2) for valid
// chip pixels, divided by the image's maximum; then every canvas pixel is owned by the image with the
// largest normalised distance (strict >, so the lowest image index wins ties and 0 never wins) and only
// the owner's mask is 255.
// Here no distance map is ever stored (the reference keeps an f32 map per chip):
//   k6_dist_max   per chip pixel, compute only: validity from the warp's own coordinate chain (M/MosaicImage.cpp:2356-2362,
//                 :2370), distance to the nearest edge, per-image maximum (atomicMax on the float bits);
//   k6_owner      per CANVAS pixel: the normalised distance of every chip whose box covers the pixel is evaluated once, in
//                 image index order with the reference's strict >; then each of those chips gets its mask byte (255 for the
//                 owner, 0 for the others) and the bounding box of the pixels it owns (K7 only builds pyramids there).
// The arithmetic per (chip, pixel) is exactly the reference's: same expressions, same order, IEEE division by the maximum.
#include <math.h>
#include <string.h>
#include "canvas.h"

namespace {

constexpr int kTileW = 128, kTileH = 8;

// distance of chip pixel (c, r) to the nearest quad edge; 0 when the pixel is outside the source frame (mask == 0)
__device__ __forceinline__ float chip_min_dist(const ChipDesc& D, int c, int r, float dgx, float dgy, float w1f, float h1f)
{
    const float yt = (float)r - dgy - D.sy + (float)D.beg_y;          // yTemp (:2357)
    const float ya = yt * D.inv[1], yb = yt * D.inv[4];
    const float xt = (float)c - dgx - D.sx + (float)D.beg_x;          // xTemp (:2356)
    float xs = xt * D.inv[0] + ya + D.inv[2];
    float ys = xt * D.inv[3] + yb + D.inv[5];
    if (!D.affine) {
        const float den = xt * D.inv[6] + yt * D.inv[7] + D.inv[8];
        xs = xs / den; ys = ys / den;
    }
    if (!((xs >= 0.0f) && (xs < w1f) && (ys >= 0.0f) && (ys < h1f))) return 0.0f;
    float mind = 536870912.0f;                                         // float minDist = 1<<29
#pragma unroll
    for (int e = 0; e < 4; e++) {
        const float d = fabsf(D.lineA[e] * (float)c + D.lineB[e] * (float)r + D.lineC[e]) * D.lineInv[e];
        if (d < mind) mind = d;
    }
    return mind;
}

__global__ void __launch_bounds__(256)
k6_dist_max(const ChipDesc* __restrict__ descs, float* __restrict__ dist_max, float dgx, float dgy, float w1f, float h1f)
{
    const ChipDesc& D = descs[blockIdx.z];
    if (!D.keep) return;
    if (blockIdx.x * kTileW >= D.chip_w || blockIdx.y * kTileH >= D.chip_h) return;
    const int r = blockIdx.y * kTileH + threadIdx.y;
    float mx = 0.0f;
    if (r < D.chip_h) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int c = blockIdx.x * kTileW + threadIdx.x + 32 * i;
            if (c < D.chip_w) mx = fmaxf(mx, chip_min_dist(D, c, r, dgx, dgy, w1f, h1f));
        }
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
        if (m > 0.0f && m > dist_max[blockIdx.z]) atomicMax(reinterpret_cast<unsigned int*>(dist_max + blockIdx.z), __float_as_uint(m));
    }
}

// chips [base, base + 256) whose boxes intersect the tile, compacted in index order (one candidate per thread)
__device__ __forceinline__ int tile_box_list(const ChipBox* __restrict__ box, int n, int base, int tx0, int ty0, int tx1, int ty1, int* list, int* wcount)
{
    const int tid = threadIdx.y * 32 + threadIdx.x, lane = threadIdx.x, warp = threadIdx.y;
    const int c = base + tid;
    bool hit = false;
    if (c < n) {
        const ChipBox b = box[c];
        hit = b.w > 0 && b.beg_x < tx1 && b.beg_x + b.w > tx0 && b.beg_y < ty1 && b.beg_y + b.h > ty0;
    }
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if (lane == 0) wcount[warp] = __popc(m);
    __syncthreads();
    int off = 0, total = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) { const int v = wcount[k]; if (k < warp) off += v; total += v; }
    if (hit) list[off + __popc(m & ((1u << lane) - 1u))] = c;
    __syncthreads();
    return total;
}

// canvas rectangle [rx0, rx1) x [ry0, ry1); a CTA covers 128 x 8 canvas pixels, a thread the pixels x + 32 k of one row
__global__ void __launch_bounds__(256)
k6_owner(const ChipDesc* __restrict__ descs, const ChipBox* __restrict__ box, const float* __restrict__ dist_max, int n,
         float dgx, float dgy, float w1f, float h1f, int rx0, int ry0, int rx1, int ry1, int32_t* __restrict__ own_bbox)
{
    __shared__ int list[256];
    __shared__ int wcount[8];
    __shared__ int sbb[256][4];
    const int tx0 = rx0 + blockIdx.x * kTileW, ty0 = ry0 + blockIdx.y * kTileH;
    const int gy = ty0 + threadIdx.y;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    float best[4] = {0.0f, 0.0f, 0.0f, 0.0f};                         // float maxDist = 0 (:1850)
    int owner[4] = {-1, -1, -1, -1};
    // pass 0: arg-max in image index order; pass 1: mask bytes + owned bounding boxes
    for (int pass = 0; pass < 2; pass++) {
        for (int base = 0; base < n; base += 256) {
            const int total = tile_box_list(box, n, base, tx0, ty0, min(tx0 + kTileW, rx1), min(ty0 + kTileH, ry1), list, wcount);
            if (pass == 1) {
                if (tid < total) { sbb[tid][0] = 1 << 30; sbb[tid][1] = 1 << 30; sbb[tid][2] = -1; sbb[tid][3] = -1; }
                __syncthreads();
            }
            for (int e = 0; e < total; e++) {
                const int m = list[e];
                const ChipDesc& D = descs[m];
                const int r = gy - D.beg_y;
                if (gy >= ry1 || r < 0 || r >= D.chip_h) continue;
                if (pass == 0) {
                    const float mx = dist_max[m];
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const int gx = tx0 + threadIdx.x + 32 * k, c = gx - D.beg_x;
                        if (gx >= rx1 || c < 0 || c >= D.chip_w) continue;
                        const float md = chip_min_dist(D, c, r, dgx, dgy, w1f, h1f);
                        if (md == 0.0f) continue;                         // invalid pixel: the map holds 0 there
                        const float dn = md / mx;                         // pMapRow[c] /= maxDist (:1830)
                        if (dn > best[k]) { best[k] = dn; owner[k] = m; }
                    }
                } else {
                    uint8_t* mrow = D.mask + (size_t)r * D.mask_step;
                    int bx0 = 1 << 30, bx1 = -1;
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const int gx = tx0 + threadIdx.x + 32 * k, c = gx - D.beg_x;
                        if (gx >= rx1 || c < 0 || c >= D.chip_w) continue;
                        const bool mine = owner[k] == m;
                        mrow[c] = mine ? 255 : 0;
                        if (mine) { bx0 = min(bx0, c); bx1 = max(bx1, c); }
                    }
                    const int wx0 = __reduce_min_sync(0xffffffffu, bx0), wx1 = __reduce_max_sync(0xffffffffu, bx1);
                    if (threadIdx.x == 0 && wx1 >= 0) {
                        atomicMin(&sbb[e][0], wx0); atomicMax(&sbb[e][2], wx1);
                        atomicMin(&sbb[e][1], r); atomicMax(&sbb[e][3], r);
                    }
                }
            }
            __syncthreads();
            if (pass == 1 && tid < total && sbb[tid][2] >= 0) {
                int32_t* g = own_bbox + (size_t)list[tid] * 4;
                if (sbb[tid][0] < g[0]) atomicMin(g + 0, sbb[tid][0]);
                if (sbb[tid][1] < g[1]) atomicMin(g + 1, sbb[tid][1]);
                if (sbb[tid][2] > g[2]) atomicMax(g + 2, sbb[tid][2]);
                if (sbb[tid][3] > g[3]) atomicMax(g + 3, sbb[tid][3]);
            }
            __syncthreads();
        }
    }
}

__global__ void k6_init_bbox(int32_t* __restrict__ bb, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { bb[4 * i] = 1 << 30; bb[4 * i + 1] = 1 << 30; bb[4 * i + 2] = -1; bb[4 * i + 3] = -1; }
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
    if (!cv->d_dist_max) {
        UAVM_CUDA(ctx, cudaMalloc(&cv->d_dist_max, (size_t)cv->n * sizeof(float)));
        UAVM_CUDA(ctx, cudaMalloc(&cv->d_box, (size_t)cv->n * sizeof(ChipBox)));
        UAVM_CUDA(ctx, cudaMalloc(&cv->d_own_bbox, (size_t)cv->n * 4 * sizeof(int32_t)));
        cv->lines_dirty = true;
    }
    if (cv->lines_dirty) {
        // edge lines of every active chip (:1783-1787) and the compact boxes of the tile scans
        std::vector<ChipBox> boxes(cv->n);
        for (int k = 0; k < cv->n; k++) {
            ChipDesc& d = cv->desc[k];
            boxes[k] = ChipBox{d.beg_x, d.beg_y, d.keep ? d.chip_w : 0, d.keep ? d.chip_h : 0};
            if (!d.keep) continue;
            for (int e = 0; e < 4; e++) {
                const int f = (e + 1) & 3;
                line_of_2_points(d.lineA[e], d.lineB[e], d.lineC[e], d.quad[2 * e], d.quad[2 * e + 1], d.quad[2 * f], d.quad[2 * f + 1]);
                d.lineInv[e] = 1 / sqrtf(d.lineA[e] * d.lineA[e] + d.lineB[e] * d.lineB[e]);
            }
        }
        UAVM_CUDA(ctx, cudaMemcpyAsync(cv->d_box, boxes.data(), boxes.size() * sizeof(ChipBox), cudaMemcpyHostToDevice, ctx->stream));
        int rc = uavm_canvas_upload_desc(ctx, cv);          // synchronises: `boxes` may go out of scope
        if (rc != UAVM_OK) return rc;
        cv->lines_dirty = false;
    }
    UAVM_CUDA(ctx, cudaMemsetAsync(cv->d_dist_max, 0, (size_t)cv->n * sizeof(float), ctx->stream));
    k6_init_bbox<<<(cv->n + 255) / 256, 256, 0, ctx->stream>>>(cv->d_own_bbox, cv->n);
    UAVM_CHECK_LAUNCH(ctx);
    const float w1f = (float)(cv->img_w - 1), h1f = (float)(cv->img_h - 1);
    {
        dim3 grid((cv->max_chip_w + kTileW - 1) / kTileW, (cv->max_chip_h + kTileH - 1) / kTileH, cv->n);
        k6_dist_max<<<grid, dim3(32, 8), 0, ctx->stream>>>(cv->d_desc, cv->d_dist_max, cv->layout.dgx, cv->layout.dgy, w1f, h1f);
        UAVM_CHECK_LAUNCH(ctx);
    }
    {
        // canvas pixels whose masks K7 can read: the whole canvas, or this context's rectangle + kShardMargin
        int rx0 = 0, ry0 = 0, rx1 = cv->layout.canvas_w, ry1 = cv->layout.canvas_h;
        if (cv->sharded) {
            rx0 = cv->rect_x0 - uavm_canvas::kShardMargin > 0 ? cv->rect_x0 - uavm_canvas::kShardMargin : 0;
            ry0 = cv->rect_y0 - uavm_canvas::kShardMargin > 0 ? cv->rect_y0 - uavm_canvas::kShardMargin : 0;
            rx1 = cv->rect_x1 + uavm_canvas::kShardMargin < rx1 ? cv->rect_x1 + uavm_canvas::kShardMargin : rx1;
            ry1 = cv->rect_y1 + uavm_canvas::kShardMargin < ry1 ? cv->rect_y1 + uavm_canvas::kShardMargin : ry1;
        }
        dim3 grid((rx1 - rx0 + kTileW - 1) / kTileW, (ry1 - ry0 + kTileH - 1) / kTileH);
        if (grid.y > 65535) { UAVM_SET_ERR(ctx, "seam_masks: canvas rectangle too tall (%d rows): shard the canvas", ry1 - ry0); return UAVM_EINVAL; }
        k6_owner<<<grid, dim3(32, 8), 0, ctx->stream>>>(cv->d_desc, cv->d_box, cv->d_dist_max, cv->n, cv->layout.dgx, cv->layout.dgy, w1f, h1f,
                                                        rx0, ry0, rx1, ry1, cv->d_own_bbox);
        UAVM_CHECK_LAUNCH(ctx);
    }
    cv->seamed = true; cv->mask_plane_valid = true; cv->own_bbox_valid = false;
    return UAVM_OK;
}
