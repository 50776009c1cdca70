// masks.cu — K6: distance-map seam masks.  Replaces FindMasksByDistMap (M/MosaicImage.cpp:1761-1881).
//
// Reference: per image, d(x,y) = min over the 4 quad edges of |A x + B y + C| / sqrt(A^2+B^2) for valid
// chip pixels, divided by the image's maximum; then every canvas pixel is owned by the image with the
// largest normalised distance (strict >, so the lowest image index wins ties and 0 never wins) and only
// the owner's mask is 255.
// Here no distance map is ever stored (the reference keeps an f32 map per chip):
//   k6_dist_max   per chip, compute only: validity from the warp's own coordinate chain (M/MosaicImage.cpp:2356-2362, :2370),
//                 distance to the nearest edge, per-image maximum (atomicMax on the float bits); tiles that provably cannot
//                 hold the maximum are skipped (the distance is 1-Lipschitz), the maximum itself is exact;
//   k6_owner      per CANVAS pixel: the normalised distance of every chip that can own the pixel is evaluated once, in image
//                 index order with the reference's strict > (chips that provably lose everywhere in the tile are pruned with
//                 Lipschitz bounds); the owner's mask byte becomes 255 (the planes are zeroed first) and each chip gets the
//                 bounding box of the pixels it owns (K7 only builds pyramids there).
// The arithmetic per (chip, pixel) is exactly the reference's: same expressions, same order, IEEE division by the maximum.
#include <math.h>
#include <string.h>
#include "canvas.h"

namespace {

constexpr int kTileW = 128, kTileH = 8;

// Per-chip constants of the distance evaluation, packed so a thread fetches them with a few 16-byte loads and keeps them
// in registers (reading the 200-byte ChipDesc field by field per pixel made the kernels LSU / issue bound).
struct __align__(16) K6Chip {
    float inv[8];                    // inv[0..7] of the chip (inv[8] below)
    float inv8, sx, sy, fbx;         // fbx = (float)beg_x
    float fby, pad0, pad1, pad2;     // fby = (float)beg_y
    float A[4], B[4], C[4], I[4];    // quad edges A x + B y + C = 0 and 1 / sqrt(A^2 + B^2)
    int32_t beg_x, beg_y, w, h;
    int32_t affine, keep, pad3, pad4;
};

struct RowTerms { float ya, yb, yd, Br[4]; };       // row-invariant parts of the chain for one chip row

__device__ __forceinline__ RowTerms row_terms(const K6Chip& D, int r, float dgy)
{
    RowTerms t;
    const float fr = (float)r;
    const float yt = fr - dgy - D.sy + D.fby;                          // yTemp (:2357)
    t.ya = yt * D.inv[1]; t.yb = yt * D.inv[4]; t.yd = yt * D.inv[7];
#pragma unroll
    for (int e = 0; e < 4; e++) t.Br[e] = D.B[e] * fr;
    return t;
}

// distance of chip pixel (column fc = (float)c, row terms t) to the nearest quad edge; 0 when the pixel is outside the source
// frame (mask == 0).  Same expressions, same order as the reference: xTemp (:2356), the inverse map (:2359-2362), the
// validity test (:2370) and |A x + B y + C| * inv (:1789-1797).
__device__ __forceinline__ float chip_min_dist(const K6Chip& D, const RowTerms& t, float fc, float dgx, float w1f, float h1f)
{
    const float xt = fc - dgx - D.sx + D.fbx;
    float xs = xt * D.inv[0] + t.ya + D.inv[2];
    float ys = xt * D.inv[3] + t.yb + D.inv[5];
    if (!D.affine) {
        const float den = xt * D.inv[6] + t.yd + D.inv8;
        xs = xs / den; ys = ys / den;
    }
    if (!((xs >= 0.0f) && (xs < w1f) && (ys >= 0.0f) && (ys < h1f))) return 0.0f;
    float mind = 536870912.0f;                                         // float minDist = 1<<29
#pragma unroll
    for (int e = 0; e < 4; e++) {
        const float d = fabsf(D.A[e] * fc + t.Br[e] + D.C[e]) * D.I[e];
        if (d < mind) mind = d;
    }
    return mind;
}

// geometric distance to the nearest edge line, validity ignored (1-Lipschitz in the pixel position)
__device__ __forceinline__ float chip_edge_dist(const K6Chip& D, float fc, float fr)
{
    float mind = 536870912.0f;
#pragma unroll
    for (int e = 0; e < 4; e++) mind = fminf(mind, fabsf(D.A[e] * fc + D.B[e] * fr + D.C[e]) * D.I[e]);
    return mind;
}

__device__ __forceinline__ void block_max_to_global(float mx, float* slot)
{
    // block max -> one atomic per block (distances are >= 0, so the uint order of the bits is the float order)
    __shared__ float smax[8];
    mx = __int_as_float(__reduce_max_sync(0xffffffffu, __float_as_int(mx)));
    if (threadIdx.x == 0) smax[threadIdx.y] = mx;
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        float m = smax[0];
#pragma unroll
        for (int i = 1; i < 8; i++) m = fmaxf(m, smax[i]);
        if (m > 0.0f && m > *slot) atomicMax(reinterpret_cast<unsigned int*>(slot), __float_as_uint(m));
    }
}

// Per-image maximum of the distance map, two launches.  COARSE: every 8th pixel of every 8th row — a true value at a true
// pixel, so a lower bound L of the maximum.  Fine: the distance is 1-Lipschitz in the pixel position, so a 128 x 32 tile whose
// centre distance + half diagonal (+ 1 px of slack for float rounding) is below L cannot hold the maximum and is skipped;
// the tiles along the quad's medial axis are evaluated at every pixel.  The result is the exact maximum over all pixels.
constexpr int kMaxTileH = 32;
template <bool COARSE>
__global__ void __launch_bounds__(256)
k6_dist_max(const K6Chip* __restrict__ chips, float* __restrict__ dist_max, float dgx, float dgy, float w1f, float h1f)
{
    const K6Chip D = chips[blockIdx.z];
    if (!D.keep) return;
    float mx = 0.0f;
    if (COARSE) {
        const int c = 8 * (blockIdx.x * 32 + threadIdx.x), r = 8 * (blockIdx.y * 8 + threadIdx.y);
        if (8 * blockIdx.x * 32 >= D.w || 8 * blockIdx.y * 8 >= D.h) return;
        if (c < D.w && r < D.h) mx = chip_min_dist(D, row_terms(D, r, dgy), (float)c, dgx, w1f, h1f);
    } else {
        const int c0 = blockIdx.x * kTileW, r0 = blockIdx.y * kMaxTileH;
        if (c0 >= D.w || r0 >= D.h) return;
        const float ub = chip_edge_dist(D, (float)c0 + 63.5f, (float)r0 + 15.5f) + 67.0f;      // half diagonal of 128 x 32 = 65.97
        if (ub < dist_max[blockIdx.z]) return;
#pragma unroll 1
        for (int j = 0; j < kMaxTileH / 8; j++) {
            const int r = r0 + 8 * j + threadIdx.y;
            if (r >= D.h) break;
            const RowTerms t = row_terms(D, r, dgy);
            const float fc0 = (float)(c0 + threadIdx.x);
#pragma unroll
            for (int i = 0; i < 4; i++)
                if (c0 + threadIdx.x + 32 * i < D.w) mx = fmaxf(mx, chip_min_dist(D, t, fc0 + 32.0f * i, dgx, w1f, h1f));
        }
    }
    block_max_to_global(mx, dist_max + blockIdx.z);
}

// Candidate chips of a canvas tile, for chips [base, base + 256): one chip per thread.  A chip is a candidate when its box
// intersects the tile AND it can own a pixel of it.  The second test prunes with bounds that hold for every pixel of the tile:
// the distance to the nearest edge line is 1-Lipschitz, so with R = half diagonal of the tile (+ slack)
//     dn_m(p) <= (d_m(centre) + R) / max_m                                   for every chip m
//     dn_b(p) >= (d_b(centre) - R) / max_b   if the disc of radius R + 2 around the centre is inside chip b's quad
//                                            (then every pixel of the tile is valid in b)
// A chip whose upper bound is below the best lower bound is never the arg-max inside the tile; the survivors are evaluated
// exactly, in index order.  `lbmax` carries the best lower bound over the rounds.  All 256 threads must call this.
__device__ __forceinline__ int tile_candidates(const K6Chip* __restrict__ chips, const float* __restrict__ dist_max, int n, int base,
                                               int tx0, int ty0, int tx1, int ty1, float dgx, float dgy, float w1f, float h1f,
                                               float& lbmax, int* list, int* wcount, float* s_lb)
{
    const int tid = threadIdx.y * 32 + threadIdx.x, lane = threadIdx.x, warp = threadIdx.y;
    const int c = base + tid;
    bool hit = false;
    float ub = 0.0f, lb = 0.0f;
    if (c < n) {
        const int4 b = *reinterpret_cast<const int4*>(&chips[c].beg_x);
        hit = chips[c].keep && b.x < tx1 && b.x + b.z > tx0 && b.y < ty1 && b.y + b.w > ty0;
        if (hit) {
            const K6Chip D = chips[c];
            const float mx = dist_max[c];
            const int pcx = (tx0 + tx1) >> 1, pcy = (ty0 + ty1) >> 1;                      // centre pixel of the tile
            const float R = 0.5f * sqrtf((float)((tx1 - tx0) * (tx1 - tx0) + (ty1 - ty0) * (ty1 - ty0))) + 2.0f;
            const float fc = (float)(pcx - D.beg_x), fr = (float)(pcy - D.beg_y);
            const float dc = chip_edge_dist(D, fc, fr);
            if (!(mx > 0.0f)) hit = false;                                                  // no valid pixel at all: never an owner
            else {
                ub = (dc + R) / mx * 1.0001f;
                const bool centre_valid = pcx >= D.beg_x && pcx < D.beg_x + D.w && pcy >= D.beg_y && pcy < D.beg_y + D.h &&
                                          chip_min_dist(D, row_terms(D, pcy - D.beg_y, dgy), fc, dgx, w1f, h1f) > 0.0f;
                if (centre_valid && dc > R + 2.0f) lb = (dc - R) / mx * 0.9999f;
            }
        }
    }
    // block maximum of the lower bounds
    float wl = __int_as_float(__reduce_max_sync(0xffffffffu, __float_as_int(lb)));          // lb >= 0: int order == float order
    if (lane == 0) s_lb[warp] = wl;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 8; k++) lbmax = fmaxf(lbmax, s_lb[k]);
    hit = hit && ub >= lbmax;
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

// canvas rectangle [rx0, rx1) x [ry0, ry1); a CTA covers 128 x 8 canvas pixels, a thread the pixels x + 32 k of one row.
// The mask planes are zeroed beforehand; the kernel stores 255 at the owner's pixel only.
__global__ void __launch_bounds__(256)
k6_owner(const K6Chip* __restrict__ chips, uint8_t* const* __restrict__ mask_ptr, const int32_t* __restrict__ mask_step,
         const float* __restrict__ dist_max, int n, float dgx, float dgy, float w1f, float h1f, int rx0, int ry0, int rx1, int ry1,
         int32_t* __restrict__ own_bbox)
{
    __shared__ int list[256];
    __shared__ int wcount[8];
    __shared__ float s_lb[8];
    const int tx0 = rx0 + blockIdx.x * kTileW, ty0 = ry0 + blockIdx.y * kTileH;
    const int tx1 = min(tx0 + kTileW, rx1), ty1 = min(ty0 + kTileH, ry1);
    const int gy = ty0 + threadIdx.y;
    const int gx0 = tx0 + threadIdx.x;
    float best[4] = {0.0f, 0.0f, 0.0f, 0.0f};                         // float maxDist = 0 (:1850)
    int owner[4] = {-1, -1, -1, -1};
    float lbmax = 0.0f;
    for (int base = 0; base < n; base += 256) {
        const int total = tile_candidates(chips, dist_max, n, base, tx0, ty0, tx1, ty1, dgx, dgy, w1f, h1f, lbmax, list, wcount, s_lb);
        for (int e = 0; e < total; e++) {
            const int m = list[e];
            const int4 bx = *reinterpret_cast<const int4*>(&chips[m].beg_x);
            const int r = gy - bx.y;
            if (gy >= ry1 || r < 0 || r >= bx.w) continue;                                // warp uniform (a warp = one canvas row)
            const K6Chip D = chips[m];
            const RowTerms t = row_terms(D, r, dgy);
            const float mx = dist_max[m];
            // the IEEE division is only needed when the quotient can exceed the running maximum: md / mx > best
            // implies md > best * mx * (1 - 2^-22); the guard below is far more conservative than that
            const float mxg = mx * 0.99999f;
            const float fc0 = (float)(gx0 - D.beg_x);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int gx = gx0 + 32 * k, c = gx - D.beg_x;
                if (gx >= rx1 || c < 0 || c >= D.w) continue;
                const float md = chip_min_dist(D, t, fc0 + 32.0f * k, dgx, w1f, h1f);
                if (md == 0.0f || md < best[k] * mxg) continue;           // invalid pixel (the map holds 0 there), or cannot win
                const float dn = md / mx;                                 // pMapRow[c] /= maxDist (:1830)
                if (dn > best[k]) { best[k] = dn; owner[k] = m; }
            }
        }
        __syncthreads();
    }
    // the owner's mask byte, and the bounding box of what each chip owns (lanes with the same owner reduce together)
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int m = owner[k];
        int c = 0, r = 0;
        if (m >= 0) {
            const int4 bx = *reinterpret_cast<const int4*>(&chips[m].beg_x);
            c = gx0 + 32 * k - bx.x; r = gy - bx.y;
            mask_ptr[m][(size_t)r * mask_step[m] + c] = 255;
        }
        const unsigned grp = __match_any_sync(0xffffffffu, m);
        const int c0 = __reduce_min_sync(grp, c), c1 = __reduce_max_sync(grp, c);
        if (m >= 0 && threadIdx.x == (unsigned)(__ffs(grp) - 1)) {
            int32_t* g = own_bbox + (size_t)m * 4;
            if (c0 < g[0]) atomicMin(g + 0, c0);
            if (r < g[1]) atomicMin(g + 1, r);
            if (c1 > g[2]) atomicMax(g + 2, c1);
            if (r > g[3]) atomicMax(g + 3, r);
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
    // the chips are not needed: validity comes from the warp's coordinate chain, so K6 may run before K5 (uavm_canvas_warp_for_blend)
    if (cv->max_chip_w <= 0) return UAVM_OK;
    UAVM_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!cv->d_dist_max) {
        UAVM_CUDA(ctx, cudaMalloc(&cv->d_dist_max, (size_t)cv->n * sizeof(float)));
        UAVM_CUDA(ctx, cudaMalloc(&cv->d_k6, (size_t)cv->n * sizeof(K6Chip)));
        UAVM_CUDA(ctx, cudaMalloc(&cv->d_mask_ptr, (size_t)cv->n * sizeof(uint8_t*)));
        UAVM_CUDA(ctx, cudaMalloc(&cv->d_mask_step, (size_t)cv->n * sizeof(int32_t)));
        UAVM_CUDA(ctx, cudaMalloc(&cv->d_own_bbox, (size_t)cv->n * 4 * sizeof(int32_t)));
        cv->lines_dirty = true;
    }
    if (cv->lines_dirty) {
        // edge lines of every active chip (:1783-1787) and the packed per-chip constants
        std::vector<K6Chip> kc(cv->n);
        std::vector<uint8_t*> mp(cv->n); std::vector<int32_t> ms(cv->n);
        for (int k = 0; k < cv->n; k++) {
            ChipDesc& d = cv->desc[k];
            K6Chip& q = kc[k]; memset(&q, 0, sizeof(q));
            mp[k] = d.mask; ms[k] = d.mask_step;
            q.beg_x = d.beg_x; q.beg_y = d.beg_y; q.w = d.keep ? d.chip_w : 0; q.h = d.keep ? d.chip_h : 0; q.keep = d.keep; q.affine = d.affine;
            if (!d.keep) continue;
            for (int e = 0; e < 4; e++) {
                const int f = (e + 1) & 3;
                line_of_2_points(d.lineA[e], d.lineB[e], d.lineC[e], d.quad[2 * e], d.quad[2 * e + 1], d.quad[2 * f], d.quad[2 * f + 1]);
                d.lineInv[e] = 1 / sqrtf(d.lineA[e] * d.lineA[e] + d.lineB[e] * d.lineB[e]);
                q.A[e] = d.lineA[e]; q.B[e] = d.lineB[e]; q.C[e] = d.lineC[e]; q.I[e] = d.lineInv[e];
            }
            memcpy(q.inv, d.inv, 8 * sizeof(float)); q.inv8 = d.inv[8];
            q.sx = d.sx; q.sy = d.sy; q.fbx = (float)d.beg_x; q.fby = (float)d.beg_y;
        }
        UAVM_CUDA(ctx, cudaMemcpyAsync(cv->d_k6, kc.data(), kc.size() * sizeof(K6Chip), cudaMemcpyHostToDevice, ctx->stream));
        UAVM_CUDA(ctx, cudaMemcpyAsync(cv->d_mask_ptr, mp.data(), mp.size() * sizeof(uint8_t*), cudaMemcpyHostToDevice, ctx->stream));
        UAVM_CUDA(ctx, cudaMemcpyAsync(cv->d_mask_step, ms.data(), ms.size() * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
        UAVM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));          // the host vectors go out of scope
        cv->lines_dirty = false;
    }
    const K6Chip* k6 = (const K6Chip*)cv->d_k6;
    UAVM_CUDA(ctx, cudaMemsetAsync(cv->d_dist_max, 0, (size_t)cv->n * sizeof(float), ctx->stream));
    UAVM_CUDA(ctx, cudaMemsetAsync(cv->d_masks, 0, cv->masks_bytes, ctx->stream));          // k6_owner stores 255 at owned pixels only
    k6_init_bbox<<<(cv->n + 255) / 256, 256, 0, ctx->stream>>>(cv->d_own_bbox, cv->n);
    UAVM_CHECK_LAUNCH(ctx);
    const float w1f = (float)(cv->img_w - 1), h1f = (float)(cv->img_h - 1);
    {
        dim3 gc((cv->max_chip_w + 255) / 256, (cv->max_chip_h + 63) / 64, cv->n);
        k6_dist_max<true><<<gc, dim3(32, 8), 0, ctx->stream>>>(k6, cv->d_dist_max, cv->layout.dgx, cv->layout.dgy, w1f, h1f);
        UAVM_CHECK_LAUNCH(ctx);
        dim3 grid((cv->max_chip_w + kTileW - 1) / kTileW, (cv->max_chip_h + kMaxTileH - 1) / kMaxTileH, cv->n);
        k6_dist_max<false><<<grid, dim3(32, 8), 0, ctx->stream>>>(k6, cv->d_dist_max, cv->layout.dgx, cv->layout.dgy, w1f, h1f);
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
        k6_owner<<<grid, dim3(32, 8), 0, ctx->stream>>>(k6, cv->d_mask_ptr, cv->d_mask_step, cv->d_dist_max, cv->n, cv->layout.dgx, cv->layout.dgy, w1f, h1f,
                                                        rx0, ry0, rx1, ry1, cv->d_own_bbox);
        UAVM_CHECK_LAUNCH(ctx);
    }
    cv->seamed = true; cv->mask_plane_valid = true; cv->own_bbox_valid = false;
    return UAVM_OK;
}
