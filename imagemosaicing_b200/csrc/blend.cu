// blend.cu — K7: multi-band blend.  Replaces detail::MultiBandBlender prepare / feed / blend and the final
// convertTo(CV_8U) as driven by LaplacianPyramidBlending (M/MosaicImage.cpp:2296-2299, :2476-2486).
//
// The blender itself is OpenCV code (third party, not under /root/reference); the arithmetic restated here is
// the one documented in oracle/oracle_blend.c: int16 Laplacian pyramids ([1 4 6 4 1] pyrDown with
// (sum+128)>>8, pyrUp with (sum+32)>>6, reflect-101 / replicate borders), f32 Gaussian weight pyramids in OpenCV's
// operation order, dst += short(src * w), wsum += w, dst = short(dst / (wsum + 1e-5)), collapse with saturating adds.
//
// Organisation (B200-first; OpenCV feeds image after image and keeps a full canvas pyramid of int16 x3 + f32):
//   1. per chip, all chips in one launch per level (blockIdx.z = chip): Gaussian pyramids of the chip's ROI, restricted to
//      the rectangles C_i the planner derives from the chip's seam-mask bounding box (blend_plan.h).  Level 0 is never
//      materialised: the first pyrDown reads the BGRA chip and its mask directly (copyMakeBorder(REFLECT) / zero border).
//   2. per canvas level, top down, ONE kernel over canvas tiles: gather the weighted Laplacian contributions of the chips
//      that touch the tile IN IMAGE INDEX ORDER (the f32 weight sums are order dependent; the int16 sums wrap and are
//      not), normalise, add pyrUp of the already final coarser level, store the final level (int16 x4) — or, at level 0,
//      the cropped u8 mosaic and its mask.  The canvas Laplacian / weight pyramids are never stored.
// A context that owns only a rectangle of the canvas (uavm_canvas_set_rect) runs the same kernels on the level rectangles
// S_i its output depends on; chip pyramids are computed in ROI coordinates exactly as in the unsharded case, so sharded
// results are bit-identical by construction (no halo heuristics).
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "canvas.h"
#include "blend_plan.h"

namespace {

using namespace uavm_plan;

struct BlendChip {                          // device-visible plan of one active chip
    const uint32_t* chip; const uint8_t* mask;
    int32_t chip_step, mask_step, cw, ch, left, top;
    int32_t tlx, tly;                       // ROI origin in padded canvas coordinates (multiples of 2^bands)
    int32_t pw[kMaxBands + 1], ph[kMaxBands + 1];
    int32_t ux0[kMaxBands + 1], uy0[kMaxBands + 1], ux1[kMaxBands + 1], uy1[kMaxBands + 1];
    int32_t cx0[kMaxBands + 1], cy0[kMaxBands + 1], cw_[kMaxBands + 1], ch_[kMaxBands + 1];
    uint32_t* pyr[kMaxBands + 1];           // levels >= 1: Gaussian levels of C_i as B | G << 8 | R << 16 words (a Gaussian pyramid of
    float* wp[kMaxBands + 1];               // u8 pixels stays in 0..255: (sum + 128) >> 8 of 256 weights), pitch cw_[i]
};

struct LevelArgs {
    int32_t level, n_chips;
    int32_t sx0, sy0, sx1, sy1;             // S_i: canvas level-i rectangle to produce
    short* fin; int32_t fin_x0, fin_y0, fin_pitch;                               // final level i (int16 x4), storage origin = S_i origin
    const short* nxt; int32_t nxt_x0, nxt_y0, nxt_pitch, nxt_rows, nxt_w, nxt_h;           // final level i+1 storage; nxt_w/h = full level size (border rules)
    uint8_t* out; uint8_t* out_mask; int32_t cw, ch;                             // level 0: mosaic (pitch cw pixels) and its mask
    uint8_t* out_peer;                                                           // level 0: the root's mosaic over NVLink (uavm_canvas_bind_root), or null
    int32_t ox0, oy0, ox1, oy1;                                                  // level 0: output rectangle
};

struct BlendWs {
    int nb = -1;
    BlendChip* d_chips = nullptr; size_t chips_cap = 0;
    uint8_t* d_scratch = nullptr; size_t scratch_cap = 0;      // chip pyramids
    short* d_fin[kMaxBands + 1] = {nullptr}; size_t fin_cap[kMaxBands + 1] = {0};
};

__device__ __forceinline__ int reflect101(int p, int n) {
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = (p < 0) ? -p : 2 * n - 2 - p;
    return p;
}
__device__ __forceinline__ int reflect_edge(int p, int n) {     // BORDER_REFLECT
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = (p < 0) ? -p - 1 : 2 * n - 1 - p;
    return p;
}
__device__ __forceinline__ int sat16(int v) { return max(-32768, min(32767, v)); }

// Pyramid pixels are 4 x int16 (B, G, R, 0): one aligned 8-byte load / store per pixel
__device__ __forceinline__ void unpack3(int2 v, int& a, int& b, int& c) { a = (short)(v.x & 0xffff); b = v.x >> 16; c = (short)(v.y & 0xffff); }
__device__ __forceinline__ int2 pack3(int a, int b, int c) { return make_int2((a & 0xffff) | (b << 16), c & 0xffff); }
// Chip-side arithmetic is SIMD within a register: a pixel is expanded to (B | G << 16, R), two u8 channels as 16-bit lanes.
// Every sum of the 5 x 5 binomial kernel (<= 256 * 255 = 65 280) and of pyrUp (<= 64 * 255 = 16 320) fits its lane, so lanes
// never carry into each other and the packed results equal OpenCV's int arithmetic on CV_16S data bit for bit.
__device__ __forceinline__ int2 px_expand(uint32_t s) { return make_int2((int)__byte_perm(s, 0u, 0x4140), (int)__byte_perm(s, 0u, 0x4442)); }

// ---------------------------------------------------------------------------------------------------------------------
// Tiled pyrDown of one pyramid level for every active chip (blockIdx.z).  A CTA produces a 64 x 8 tile of C_{l+1} from a
// 131 x 19 source window staged in shared memory, split into even / odd columns so the stride-2 taps are unit-stride,
// conflict-free LDS; a thread filters 7 window rows horizontally once for two vertically adjacent outputs.
// L0: the source level is the chip itself — copyMakeBorder(REFLECT) of the BGRA chip (u8 -> s16) and mask / 255 with a
// zero border, evaluated on the fly (MultiBandBlender::feed).
// f32 weights follow OpenCV's operation order (oracle_blend.c: vector form for columns [1, hvec_end) of the row pass and
// [0, vvec_end) of the column pass, scalar form elsewhere), in ROI-level coordinates.
constexpr int kPdW = 64, kPdH = 8, kPdCols = 2 * kPdW + 4, kPdRows = 2 * kPdH + 3;
constexpr int kPdPairs = kPdCols / 2, kPdItems = kPdRows * kPdPairs, kPdPerThread = (kPdItems + 255) / 256;

// interior staging: the window's (even, odd) pixel pairs are spread evenly over the 256 threads (item = row * 66 + pair), all
// global loads of a thread are issued before the first conversion, and a pair lands in shared memory as one 16-byte store
// (expanded pixels) + one 8-byte store (weights).  EVEN: the pair is 8-byte aligned in the source (one load).
template <bool L0, bool EVEN>
__device__ __forceinline__ void pyrdown_stage_interior(const BlendChip& B, int sl, int sx0, int sy0, int tid, int2 (*sP)[kPdCols], float (*sW)[kPdCols])
{
    uint32_t c0[kPdPerThread], c1[kPdPerThread], m[kPdPerThread];
    float f0[kPdPerThread], f1[kPdPerThread];
    const int scx0 = L0 ? 0 : B.cx0[sl], scy0 = L0 ? 0 : B.cy0[sl], scw = L0 ? 0 : B.cw_[sl];
#pragma unroll
    for (int k = 0; k < kPdPerThread; k++) {
        const int e = tid + 256 * k;
        if (e >= kPdItems) break;
        const int r = e / kPdPairs, j = e - r * kPdPairs;
        if (L0) {
            const uint32_t* cp = B.chip + (ptrdiff_t)(sy0 + r - B.top) * B.chip_step + (sx0 - B.left + 2 * j);
            const uint8_t* mp = B.mask + (ptrdiff_t)(sy0 + r - B.top) * B.mask_step + (sx0 - B.left + 2 * j);
            if (EVEN) { const uint2 c = __ldg(reinterpret_cast<const uint2*>(cp)); c0[k] = c.x; c1[k] = c.y; m[k] = __ldg(reinterpret_cast<const unsigned short*>(mp)); }
            else { c0[k] = __ldg(cp); c1[k] = __ldg(cp + 1); m[k] = (uint32_t)__ldg(mp) | ((uint32_t)__ldg(mp + 1) << 8); }
        } else {
            const ptrdiff_t g = (ptrdiff_t)(sy0 + r - scy0) * scw + (sx0 - scx0 + 2 * j);
            const uint2 c = __ldg(reinterpret_cast<const uint2*>(B.pyr[sl] + g)); c0[k] = c.x; c1[k] = c.y;
            const float2 f = __ldg(reinterpret_cast<const float2*>(B.wp[sl] + g)); f0[k] = f.x; f1[k] = f.y;
        }
    }
#pragma unroll
    for (int k = 0; k < kPdPerThread; k++) {
        const int e = tid + 256 * k;
        if (e >= kPdItems) break;
        const int r = e / kPdPairs, j = e - r * kPdPairs;
        if (L0) { f0[k] = (float)(m[k] & 0xffu) * (float)(1. / 255.); f1[k] = (float)(m[k] >> 8) * (float)(1. / 255.); }
        const int2 a = px_expand(c0[k]), b = px_expand(c1[k]);
        *reinterpret_cast<int4*>(&sP[r][2 * j]) = make_int4(a.x, a.y, b.x, b.y);
        *reinterpret_cast<float2*>(&sW[r][2 * j]) = make_float2(f0[k], f1[k]);
    }
}

template <bool L0>
__global__ void __launch_bounds__(256)
k7_pyrdown(const BlendChip* __restrict__ chips, int sl)
{
    __shared__ __align__(16) int2 sP[kPdRows][kPdCols];                           // expanded pixels (B | G << 16, R), window column order
    __shared__ __align__(16) float sW[kPdRows][kPdCols];
    const BlendChip& B = chips[blockIdx.z];
    const int dl = sl + 1;
    const int dcw = B.cw_[dl], dch = B.ch_[dl];
    const int tx0 = blockIdx.x * kPdW, ty0 = blockIdx.y * kPdH;                 // tile origin inside C_dl
    if (tx0 >= dcw || ty0 >= dch) return;
    const int ox0 = B.cx0[dl] + tx0, oy0 = B.cy0[dl] + ty0;                      // ROI level-dl coordinates
    const int w = B.pw[sl], h = B.ph[sl];
    const int sx0 = 2 * ox0 - 2, sy0 = 2 * oy0 - 2;                              // window origin (even column)
    const int tid = threadIdx.y * kPdW + threadIdx.x;
    // source storage (levels >= 1): C_sl, origin (scx0, scy0), pitch scw; coordinates are clamped into it (window positions
    // that only feed outputs outside C_dl may fall outside C_sl)
    const int scx0 = L0 ? 0 : B.cx0[sl], scy0 = L0 ? 0 : B.cy0[sl], scw = L0 ? 0 : B.cw_[sl], sch = L0 ? 0 : B.ch_[sl];
    const uint32_t* __restrict__ src = L0 ? nullptr : B.pyr[sl];
    const float* __restrict__ wsrc = L0 ? nullptr : B.wp[sl];
    // fast path: the whole window lies inside the chip (L0) resp. inside C_sl and the ROI level (hence no border arithmetic)
    const bool interior = L0 ? (sx0 - B.left >= 0 && sx0 + kPdCols - B.left <= B.cw && sy0 - B.top >= 0 && sy0 + kPdRows - B.top <= B.ch)
                             : (sx0 >= 0 && sx0 + kPdCols <= w && sy0 >= 0 && sy0 + kPdRows <= h &&
                                sx0 - scx0 >= 0 && sx0 + kPdCols - scx0 <= scw && sy0 - scy0 >= 0 && sy0 + kPdRows - scy0 <= sch);
    if (interior) {
        if (!L0 || (B.left & 1) == 0) pyrdown_stage_interior<L0, true>(B, sl, sx0, sy0, tid, sP, sW);     // C origins below the top level are even
        else pyrdown_stage_interior<L0, false>(B, sl, sx0, sy0, tid, sP, sW);
    } else
    for (int e = tid; e < kPdItems; e += 256) {
        const int r = e / kPdPairs, j = e - r * kPdPairs;
        const int Y = reflect101(sy0 + r, h);
        const int gx = sx0 + 2 * j;
#pragma unroll
        for (int o = 0; o < 2; o++) {
            const int X = reflect101(gx + o, w);
            uint32_t c; float wv = 0.0f;
            if (L0) {
                const int iy = Y - B.top, sy = reflect_edge(iy, B.ch);
                const int ix = X - B.left, sx = reflect_edge(ix, B.cw);
                c = __ldg(B.chip + (size_t)sy * B.chip_step + sx);
                if (iy >= 0 && iy < B.ch && ix >= 0 && ix < B.cw) wv = (float)__ldg(B.mask + (size_t)iy * B.mask_step + ix) * (float)(1. / 255.);
            } else {
                const size_t g = (size_t)min(max(Y - scy0, 0), sch - 1) * scw + min(max(X - scx0, 0), scw - 1);
                c = __ldg(src + g); wv = __ldg(wsrc + g);
            }
            sP[r][2 * j + o] = px_expand(c); sW[r][2 * j + o] = wv;
        }
    }
    __syncthreads();
    const int lx = tx0 + threadIdx.x;                                           // column inside C_dl
    if (lx >= dcw) return;
    const int x = ox0 + threadIdx.x, ty = threadIdx.y;                           // outputs (x, oy0 + 2 ty) and (x, oy0 + 2 ty + 1)
    const int i = threadIdx.x;                                                   // window columns 2 i .. 2 i + 4
    const int dw = B.pw[dl];
    int width0 = (w - 3) / 2 + 1; if (w < 3) width0 = 0; if (width0 > dw) width0 = dw;
    const bool hvec = x >= 1 && x < 1 + 4 * ((width0 - 1 > 0 ? width0 - 1 : 0) / 4);
    const bool vvec = x < 4 * (dw / 4);
    const unsigned act = __activemask();
    const unsigned hv = __ballot_sync(act, hvec);
    const int hmode = hv == act ? 1 : (hv == 0u ? 0 : 2);                       // warp-uniform: one form per warp but at the borders
    uint32_t ax[2] = {0u, 0u}, ay[2] = {0u, 0u};                                // packed accumulators of the two outputs
    float F[7];
    const uint32_t kw[5] = {1u, 4u, 6u, 4u, 1u};
#pragma unroll
    for (int r = 0; r < 7; r++) {
        const int wr = 4 * ty + r;
        // five consecutive window pixels = 40 bytes at a 16-byte aligned address: two 16-byte loads + one 8-byte load
        const int4 q0 = *reinterpret_cast<const int4*>(&sP[wr][2 * i]), q1 = *reinterpret_cast<const int4*>(&sP[wr][2 * i + 2]);
        const int2 p4 = sP[wr][2 * i + 4];
        const uint32_t hx = (uint32_t)q0.x + (uint32_t)p4.x + 4u * ((uint32_t)q0.z + (uint32_t)q1.z) + 6u * (uint32_t)q1.x;
        const uint32_t hy = (uint32_t)q0.y + (uint32_t)p4.y + 4u * ((uint32_t)q0.w + (uint32_t)q1.w) + 6u * (uint32_t)q1.y;
        if (r < 5) { ax[0] += kw[r] * hx; ay[0] += kw[r] * hy; }
        if (r >= 2) { ax[1] += kw[r - 2] * hx; ay[1] += kw[r - 2] * hy; }
        const float2 w01 = *reinterpret_cast<const float2*>(&sW[wr][2 * i]), w23 = *reinterpret_cast<const float2*>(&sW[wr][2 * i + 2]);
        const float s0 = w01.x, s1 = w01.y, s2 = w23.x, s3 = w23.y, s4 = sW[wr][2 * i + 4];
        const float m6 = s2 * 6.0f, m4 = (s1 + s3) * 4.0f;
        if (hmode == 1) F[r] = m6 + (m4 + (s0 + s4));
        else if (hmode == 0) F[r] = m6 + m4 + s0 + s4;
        else F[r] = hvec ? m6 + (m4 + (s0 + s4)) : m6 + m4 + s0 + s4;
    }
    uint32_t* __restrict__ dst = B.pyr[dl];
    float* __restrict__ wdst = B.wp[dl];
#pragma unroll
    for (int o = 0; o < 2; o++) {
        const int ly = ty0 + 2 * ty + o;
        if (ly >= dch) break;
        const size_t di = (size_t)ly * dcw + lx;
        const uint32_t bg = ((ax[o] + 0x00800080u) >> 8) & 0x00ff00ffu, rr = ((ay[o] + 128u) >> 8) & 0xffu;      // (sum + 128) >> 8 per lane
        dst[di] = __byte_perm(bg, rr, 0x7420);
        const float r0 = F[2 * o], r1 = F[1 + 2 * o], r2 = F[2 + 2 * o], r3 = F[3 + 2 * o], r4 = F[4 + 2 * o];
        const float v = vvec ? ((r1 + r3) + r2) * 4.0f + ((r0 + r4) + (r2 + r2)) : r2 * 6.0f + (r1 + r3) * 4.0f + r0 + r4;
        wdst[di] = v * (1.0f / 256.0f);
    }
}

// pyrUp of a chip's Gaussian level (u8 words) for the 4 x 2 pixel block made of the quads (c0, cy) and (c0 + 1, cy), packed:
// upx[dy][px] = (B | G << 16), upy[dy][px] = R.  Per axis: even sample s[i-1] + 6 s[i] + s[i+1], odd sample 4 (s[i] + s[i+1]);
// reflect-101 at the near edge, replicate at the far edge; (sum + 32) >> 6.  lo: storage with origin (ox, oy) and `pitch`
// pixels per row; lw x lh = full level size.
__device__ __forceinline__ void pyrup_2quads_u8(const uint32_t* __restrict__ lo, int pitch, int ox, int oy, int lw, int lh, int c0, int cy,
                                                uint32_t upx[2][4], uint32_t upy[2][4])
{
    const int c1 = min(c0 + 1, lw - 1);
    const int cols[4] = {c0 == 0 ? (lw > 1 ? 1 : 0) : c0 - 1, c0, c1, c1 == lw - 1 ? lw - 1 : c1 + 1};
    const int rows[3] = {cy == 0 ? (lh > 1 ? 1 : 0) : cy - 1, cy, cy == lh - 1 ? lh - 1 : cy + 1};
    uint32_t hex[2][3], hox[2][3], hey[2][3], hoy[2][3];           // [quad][row]: even-x sum a + 6 b + c, odd-x sum 4 (b + c)
#pragma unroll
    for (int r = 0; r < 3; r++) {
        const uint32_t* rowp = lo + (ptrdiff_t)(rows[r] - oy) * pitch - ox;
        int2 v[4];
#pragma unroll
        for (int j = 0; j < 4; j++) v[j] = px_expand(__ldg(rowp + cols[j]));
        hex[0][r] = (uint32_t)v[0].x + 6u * (uint32_t)v[1].x + (uint32_t)v[2].x; hox[0][r] = 4u * ((uint32_t)v[1].x + (uint32_t)v[2].x);
        hex[1][r] = (uint32_t)v[1].x + 6u * (uint32_t)v[2].x + (uint32_t)v[3].x; hox[1][r] = 4u * ((uint32_t)v[2].x + (uint32_t)v[3].x);
        hey[0][r] = (uint32_t)v[0].y + 6u * (uint32_t)v[1].y + (uint32_t)v[2].y; hoy[0][r] = 4u * ((uint32_t)v[1].y + (uint32_t)v[2].y);
        hey[1][r] = (uint32_t)v[1].y + 6u * (uint32_t)v[2].y + (uint32_t)v[3].y; hoy[1][r] = 4u * ((uint32_t)v[2].y + (uint32_t)v[3].y);
    }
#pragma unroll
    for (int q = 0; q < 2; q++)
#pragma unroll
        for (int dx = 0; dx < 2; dx++) {
            const uint32_t x0 = dx ? hox[q][0] : hex[q][0], x1 = dx ? hox[q][1] : hex[q][1], x2 = dx ? hox[q][2] : hex[q][2];
            const uint32_t y0 = dx ? hoy[q][0] : hey[q][0], y1 = dx ? hoy[q][1] : hey[q][1], y2 = dx ? hoy[q][2] : hey[q][2];
            upx[0][2 * q + dx] = ((x0 + 6u * x1 + x2 + 0x00200020u) >> 6) & 0x03ff03ffu;
            upx[1][2 * q + dx] = ((4u * (x1 + x2) + 0x00200020u) >> 6) & 0x03ff03ffu;
            upy[0][2 * q + dx] = (y0 + 6u * y1 + y2 + 32u) >> 6;
            upy[1][2 * q + dx] = (4u * (y1 + y2) + 32u) >> 6;
        }
}

// short(d / (wsum + 1e-5f)) (normalizeUsingWeightMap), with two exact shortcuts that remove most IEEE divisions:
//   d == 0            -> 0
//   wsum == 1.0f      -> d - sign(d): d / 1.00001f lies strictly between d - sign(d) and d for every int16 d (d * 1e-5 < 1 and
//                        far above half an ulp of d), so the truncation drops exactly one unit — checked for all 65 536 values
//                        of d in tests/test_cpu_blend.py.  With seam masks every owned level-0 pixel has wsum == 1.0f.
__device__ __forceinline__ int norm_div(int d, float ws)
{
    if (d == 0) return 0;
    if (ws == 1.0f) return d - (d > 0 ? 1 : -1);
    return (short)__float2int_rz((float)d / (ws + 1e-5f));
}

// pyrUp for a 4 x 2 block from a staged coarse tile (border rules already applied while staging): st[r][c] holds the coarse
// pixel (row r0 - 1 + r, column c0 - 1 + c); lc = local column of quad 0 (>= 1), lr = local row (>= 1)
template <int PITCH>
__device__ __forceinline__ void pyrup_2quads_smem(const int2 (*st)[PITCH], int lc, int lr, int up[2][4][3])
{
    int he[2][3][3], ho[2][3][3];
#pragma unroll
    for (int r = 0; r < 3; r++) {
        int v[4][3];
#pragma unroll
        for (int j = 0; j < 4; j++) unpack3(st[lr - 1 + r][lc - 1 + j], v[j][0], v[j][1], v[j][2]);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            he[0][r][k] = v[0][k] + 6 * v[1][k] + v[2][k]; ho[0][r][k] = 4 * (v[1][k] + v[2][k]);
            he[1][r][k] = v[1][k] + 6 * v[2][k] + v[3][k]; ho[1][r][k] = 4 * (v[2][k] + v[3][k]);
        }
    }
#pragma unroll
    for (int q = 0; q < 2; q++)
#pragma unroll
        for (int dx = 0; dx < 2; dx++)
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const int h0 = dx ? ho[q][0][k] : he[q][0][k], h1 = dx ? ho[q][1][k] : he[q][1][k], h2 = dx ? ho[q][2][k] : he[q][2][k];
                up[0][2 * q + dx][k] = sat16((h0 + 6 * h1 + h2 + 32) >> 6);
                up[1][2 * q + dx][k] = sat16((4 * (h1 + h2) + 32) >> 6);
            }
}

// Sorted list of the chips whose contribution rectangle U_level intersects the CTA's tile, for chips [base, base + 256):
// one candidate per thread, compacted in index order through warp ballots.  All 256 threads must call it.
__device__ __forceinline__ int tile_chip_list(const BlendChip* __restrict__ chips, int n, int base, int level, int tx0, int ty0, int tx1, int ty1,
                                              int* list, int* wcount)
{
    const int tid = threadIdx.y * blockDim.x + threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c = base + tid;
    bool hit = false;
    if (c < n) {
        const BlendChip& B = chips[c];
        const int ofx = B.tlx >> level, ofy = B.tly >> level;
        hit = B.ux0[level] + ofx < tx1 && B.ux1[level] + ofx > tx0 && B.uy0[level] + ofy < ty1 && B.uy1[level] + ofy > ty0;
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

// ---------------------------------------------------------------------------------------------------------------------
// One canvas level below the top (it has a coarser level): a thread owns a 4 x 2 pixel block (two quads).
//   for every chip touching the tile, in index order:  lap = pyr_i - pyrUp(pyr_{i+1});  d += short(lap * w);  wsum += w
//   v = short(d / (wsum + 1e-5));  final = sat(pyrUp(final_{i+1}) + v)
// L0: pyr_0 is the chip itself (a non-zero weight implies the pixel lies inside the chip), w = mask / 255, and the result is
// written as the cropped u8 mosaic + mask (zero where wsum <= 1e-5; convertTo(CV_8U) saturates).
template <bool L0, int MINB>
__global__ void __launch_bounds__(256, MINB)
k7_level(const BlendChip* __restrict__ chips, const LevelArgs A)
{
    __shared__ int list[256];
    __shared__ int wcount[8];
    __shared__ int2 snx[10][68];                                                  // final level i+1 under this tile: 66 x 10 coarse pixels
    const int level = A.level;
    const int tx0 = (A.sx0 & ~3) + blockIdx.x * 128, ty0 = A.sy0 + blockIdx.y * 16;
    const int X = tx0 + 4 * threadIdx.x, Y = ty0 + 2 * threadIdx.y;               // this thread's block (canvas level coordinates)
    const bool live = X + 4 > A.sx0 && X < A.sx1 && Y < A.sy1;
    __shared__ int s_wide;                                                        // a staged value lies outside [-512, 511]
    // stage the coarse tile with pyrUp's border rules applied (reflect-101 at the near edge, replicate at the far edge);
    // positions outside the stored rectangle S_{i+1} only feed pixels outside S_i and are clamped into it.
    // Values are staged BIASED and packed, (B + 512 | G + 512 << 16, R + 512): while every value of the tile lies in
    // [-512, 511] (blended image data does) all pyrUp sums stay inside their 16-bit lanes (<= 64 * 1023) and
    // ((sum' + 32) >> 6) = ((sum + 32) >> 6) + 512 exactly (the bias 64 * 512 is a multiple of 64).  A tile with a wider
    // value is re-staged raw and takes the generic per-channel path.
    if (threadIdx.x == 0 && threadIdx.y == 0) s_wide = 0;
    __syncthreads();
    {
        int wide = 0;
        // tile-uniform: the 66 x 10 coarse window lies inside the level and inside the stored rectangle (no border rules, no clamps)
        const int cx_lo = (tx0 >> 1) - 1, cy_lo = (ty0 >> 1) - 1;
        const bool st_in = cx_lo >= 0 && cx_lo >= A.nxt_x0 && cx_lo + 66 <= A.nxt_w && cx_lo + 66 <= A.nxt_x0 + A.nxt_pitch &&
                           cy_lo >= 0 && cy_lo >= A.nxt_y0 && cy_lo + 10 <= A.nxt_h && cy_lo + 10 <= A.nxt_y0 + A.nxt_rows;
        for (int e = threadIdx.y * 32 + threadIdx.x; e < 10 * 66; e += 256) {
            const int r = e / 66, c = e - r * 66;
            int lx, ly;
            if (st_in) { lx = cx_lo + c - A.nxt_x0; ly = cy_lo + r - A.nxt_y0; }
            else {
                int cx = cx_lo + c, cy = cy_lo + r;
                cx = cx < 0 ? (A.nxt_w > 1 ? 1 : 0) : (cx > A.nxt_w - 1 ? A.nxt_w - 1 : cx);
                cy = cy < 0 ? (A.nxt_h > 1 ? 1 : 0) : (cy > A.nxt_h - 1 ? A.nxt_h - 1 : cy);
                lx = min(max(cx - A.nxt_x0, 0), A.nxt_pitch - 1); ly = min(max(cy - A.nxt_y0, 0), A.nxt_rows - 1);
            }
            int b3, g3, r3;
            unpack3(*reinterpret_cast<const int2*>(A.nxt + ((size_t)ly * A.nxt_pitch + lx) * 4), b3, g3, r3);
            wide |= ((unsigned)(b3 + 512) > 1023u) | ((unsigned)(g3 + 512) > 1023u) | ((unsigned)(r3 + 512) > 1023u);
            snx[r][c] = make_int2((b3 + 512) | ((g3 + 512) << 16), r3 + 512);
        }
        if (wide) s_wide = 1;
    }
    int d[2][4][3];
    float ws[2][4];
#pragma unroll
    for (int dy = 0; dy < 2; dy++)
#pragma unroll
        for (int p = 0; p < 4; p++) { ws[dy][p] = 0.0f; d[dy][p][0] = d[dy][p][1] = d[dy][p][2] = 0; }

    for (int base = 0; base < A.n_chips; base += 256) {
        const int total = tile_chip_list(chips, A.n_chips, base, level, tx0, ty0, tx0 + 128, ty0 + 16, list, wcount);
        for (int e = 0; e < total; e++) {
            const BlendChip& B = chips[list[e]];
            const int x = X - (B.tlx >> level), y = Y - (B.tly >> level);         // ROI level coordinates of the block (even)
            const int ux0 = B.ux0[level], uy0 = B.uy0[level], ux1 = B.ux1[level], uy1 = B.uy1[level];
            if (!live || x + 4 <= ux0 || x >= ux1 || y + 2 <= uy0 || y >= uy1) continue;
            float w[2][4];
            bool any = false;
            uint32_t m01[2] = {0u, 0u};
            if (L0) {
                const int ix = x - B.left, iy = y - B.top;                        // chip coordinates
#pragma unroll
                for (int dy = 0; dy < 2; dy++) {
                    const bool yin = y + dy >= uy0 && y + dy < uy1;
                    const uint8_t* mrow = B.mask + (ptrdiff_t)(iy + dy) * B.mask_step;
                    uint32_t m = 0;
                    if (yin) {
                        if (x >= ux0 && x + 4 <= ux1) {
                            // 4 mask bytes at an arbitrary byte offset: two aligned words (rows are padded to align4, + 256 B slack) and a funnel shift
                            const uint32_t* mw = reinterpret_cast<const uint32_t*>(mrow + (ix & ~3));
                            m = (ix & 3) ? __funnelshift_r(__ldg(mw), __ldg(mw + 1), 8 * (ix & 3)) : __ldg(mw);
                        } else {
#pragma unroll
                            for (int p = 0; p < 4; p++) if (x + p >= ux0 && x + p < ux1) m |= (uint32_t)__ldg(mrow + ix + p) << (8 * p);
                        }
                    }
                    m01[dy] = m;
                }
                any = (m01[0] | m01[1]) != 0u;
            } else {
                const float* __restrict__ wp = B.wp[level];
                const int lx = x - B.cx0[level], ly = y - B.cy0[level], pitch = B.cw_[level];
#pragma unroll
                for (int dy = 0; dy < 2; dy++) {
                    const bool yin = y + dy >= uy0 && y + dy < uy1;
#pragma unroll
                    for (int p = 0; p < 4; p++) {
                        float v = 0.0f;
                        if (yin && x + p >= ux0 && x + p < ux1) v = wp[(ptrdiff_t)(ly + dy) * pitch + lx + p];
                        w[dy][p] = v; any = any || v != 0.0f;
                    }
                }
            }
            if (!any) continue;                                                   // dst + short(lap * 0) == dst, wsum + 0 == wsum
            uint32_t upx[2][4], upy[2][4];                                        // pyrUp of the chip's coarser Gaussian level, packed
            pyrup_2quads_u8(B.pyr[level + 1], B.cw_[level + 1], B.cx0[level + 1], B.cy0[level + 1], B.pw[level + 1], B.ph[level + 1], x >> 1, y >> 1, upx, upy);
            if (L0) {
                if ((m01[0] & m01[1]) == 0xffffffffu) {
                    // all eight weights are exactly 1 (the interior of what the chip owns): d += pixel - pyrUp, wsum += 1, no
                    // per-pixel tests and no float round trip; four chip pixels are one 16-byte load when aligned
                    const int ix = x - B.left;
#pragma unroll
                    for (int dy = 0; dy < 2; dy++) {
                        const uint32_t* crow = B.chip + (ptrdiff_t)(y + dy - B.top) * B.chip_step + ix;
                        uint32_t sp[4];
                        if ((ix & 3) == 0) { const uint4 q = __ldg(reinterpret_cast<const uint4*>(crow)); sp[0] = q.x; sp[1] = q.y; sp[2] = q.z; sp[3] = q.w; }
                        else {
#pragma unroll
                            for (int p = 0; p < 4; p++) sp[p] = __ldg(crow + p);
                        }
#pragma unroll
                        for (int p = 0; p < 4; p++) {
                            d[dy][p][0] += (int)(sp[p] & 0xffu) - (int)(upx[dy][p] & 0xffffu);
                            d[dy][p][1] += (int)((sp[p] >> 8) & 0xffu) - (int)(upx[dy][p] >> 16);
                            d[dy][p][2] += (int)((sp[p] >> 16) & 0xffu) - (int)upy[dy][p];
                            ws[dy][p] += 1.0f;
                        }
                    }
                    continue;
                }
#pragma unroll
                for (int dy = 0; dy < 2; dy++)
#pragma unroll
                    for (int p = 0; p < 4; p++) w[dy][p] = (float)((m01[dy] >> (8 * p)) & 0xffu) * (float)(1. / 255.);
            }
#pragma unroll
            for (int dy = 0; dy < 2; dy++)
#pragma unroll
                for (int p = 0; p < 4; p++) {
                    const float wv = w[dy][p];
                    if (wv == 0.0f) continue;
                    // the chip's pixel at this level: the chip itself (level 0), else its Gaussian level; both are u8 words
                    const uint32_t s = L0 ? __ldg(B.chip + (ptrdiff_t)(y + dy - B.top) * B.chip_step + (x + p - B.left))
                                          : __ldg(B.pyr[level] + (ptrdiff_t)(y + dy - B.cy0[level]) * B.cw_[level] + (x + p - B.cx0[level]));
                    // Laplacian = pixel - pyrUp, in [-255, 255]: the saturating subtract never saturates
                    const int lap[3] = {(int)(s & 0xffu) - (int)(upx[dy][p] & 0xffffu), (int)((s >> 8) & 0xffu) - (int)(upx[dy][p] >> 16),
                                        (int)((s >> 16) & 0xffu) - (int)upy[dy][p]};
                    // dst += short(lap * w): the int16 sums wrap, so they are kept as ints and truncated once at the end;
                    // w == 1.0f (every owned level-0 pixel under seam masks) needs no float round trip
                    if (wv == 1.0f) {
#pragma unroll
                        for (int k = 0; k < 3; k++) d[dy][p][k] += lap[k];
                    } else {
#pragma unroll
                        for (int k = 0; k < 3; k++) d[dy][p][k] += (short)__float2int_rz((float)lap[k] * wv);
                    }
                    ws[dy][p] += wv;
                }
        }
        __syncthreads();
    }
    __syncthreads();                                                              // the staged tile and s_wide
    int up[2][4][3];
    if (s_wide) {                                                                 // uniform: re-stage raw values, generic per-channel pyrUp
        __syncthreads();
        for (int e = threadIdx.y * 32 + threadIdx.x; e < 10 * 66; e += 256) {
            const int r = e / 66, c = e - r * 66;
            int cx = (tx0 >> 1) - 1 + c, cy = (ty0 >> 1) - 1 + r;
            cx = cx < 0 ? (A.nxt_w > 1 ? 1 : 0) : (cx > A.nxt_w - 1 ? A.nxt_w - 1 : cx);
            cy = cy < 0 ? (A.nxt_h > 1 ? 1 : 0) : (cy > A.nxt_h - 1 ? A.nxt_h - 1 : cy);
            const int lx = min(max(cx - A.nxt_x0, 0), A.nxt_pitch - 1), ly = min(max(cy - A.nxt_y0, 0), A.nxt_rows - 1);
            snx[r][c] = *reinterpret_cast<const int2*>(A.nxt + ((size_t)ly * A.nxt_pitch + lx) * 4);
        }
        __syncthreads();
        if (!live) return;
        pyrup_2quads_smem<68>(snx, 1 + 2 * threadIdx.x, 1 + threadIdx.y, up);
    } else {
        if (!live) return;
        const int lc = 1 + 2 * threadIdx.x, lr = 1 + threadIdx.y;
        uint32_t hex[2][3], hox[2][3], hey[2][3], hoy[2][3];
#pragma unroll
        for (int r = 0; r < 3; r++) {
            const int2 v0 = snx[lr - 1 + r][lc - 1], v1 = snx[lr - 1 + r][lc], v2 = snx[lr - 1 + r][lc + 1], v3 = snx[lr - 1 + r][lc + 2];
            hex[0][r] = (uint32_t)v0.x + 6u * (uint32_t)v1.x + (uint32_t)v2.x; hox[0][r] = 4u * ((uint32_t)v1.x + (uint32_t)v2.x);
            hex[1][r] = (uint32_t)v1.x + 6u * (uint32_t)v2.x + (uint32_t)v3.x; hox[1][r] = 4u * ((uint32_t)v2.x + (uint32_t)v3.x);
            hey[0][r] = (uint32_t)v0.y + 6u * (uint32_t)v1.y + (uint32_t)v2.y; hoy[0][r] = 4u * ((uint32_t)v1.y + (uint32_t)v2.y);
            hey[1][r] = (uint32_t)v1.y + 6u * (uint32_t)v2.y + (uint32_t)v3.y; hoy[1][r] = 4u * ((uint32_t)v2.y + (uint32_t)v3.y);
        }
#pragma unroll
        for (int q = 0; q < 2; q++)
#pragma unroll
            for (int dx = 0; dx < 2; dx++) {
                const uint32_t x0 = dx ? hox[q][0] : hex[q][0], x1 = dx ? hox[q][1] : hex[q][1], x2 = dx ? hox[q][2] : hex[q][2];
                const uint32_t y0 = dx ? hoy[q][0] : hey[q][0], y1 = dx ? hoy[q][1] : hey[q][1], y2 = dx ? hoy[q][2] : hey[q][2];
                const uint32_t ex = ((x0 + 6u * x1 + x2 + 0x00200020u) >> 6) & 0x03ff03ffu, ox_ = ((4u * (x1 + x2) + 0x00200020u) >> 6) & 0x03ff03ffu;
                const uint32_t ey = (y0 + 6u * y1 + y2 + 32u) >> 6, oy_ = (4u * (y1 + y2) + 32u) >> 6;
                up[0][2 * q + dx][0] = (int)(ex & 0xffffu) - 512; up[0][2 * q + dx][1] = (int)(ex >> 16) - 512; up[0][2 * q + dx][2] = (int)ey - 512;
                up[1][2 * q + dx][0] = (int)(ox_ & 0xffffu) - 512; up[1][2 * q + dx][1] = (int)(ox_ >> 16) - 512; up[1][2 * q + dx][2] = (int)oy_ - 512;
            }
    }
    // every weight sum of the block is 0 (nothing fed: d == 0) or exactly 1 (the rule under seam masks at level 0): the
    // normalisation is d - sign(d) for all 24 values, no branches and no division
    bool unit = true;
#pragma unroll
    for (int dy = 0; dy < 2; dy++)
#pragma unroll
        for (int p = 0; p < 4; p++) unit = unit && (ws[dy][p] == 1.0f || ws[dy][p] == 0.0f);
    if (unit) {
#pragma unroll
        for (int dy = 0; dy < 2; dy++)
#pragma unroll
            for (int p = 0; p < 4; p++)
#pragma unroll
                for (int k = 0; k < 3; k++) { const int dd = (short)d[dy][p][k]; d[dy][p][k] = dd - (dd > 0) + (dd < 0); }
    } else {
#pragma unroll
        for (int dy = 0; dy < 2; dy++)
#pragma unroll
            for (int p = 0; p < 4; p++)
#pragma unroll
                for (int k = 0; k < 3; k++) d[dy][p][k] = norm_div((short)d[dy][p][k], ws[dy][p]);
    }
#pragma unroll
    for (int dy = 0; dy < 2; dy++) {
        const int yy = Y + dy;
        int v[4][3];
#pragma unroll
        for (int p = 0; p < 4; p++) {
#pragma unroll
            for (int k = 0; k < 3; k++) v[p][k] = L0 ? up[dy][p][k] + d[dy][p][k] : sat16(up[dy][p][k] + d[dy][p][k]);   // level 0 clamps to [0, 255] below
        }
        if (!L0) {
            if (yy >= A.sy1) break;
            short* frow = A.fin + ((ptrdiff_t)(yy - A.fin_y0) * A.fin_pitch - A.fin_x0) * 4;
            if (X >= A.sx0 && X + 4 <= A.sx1) {
                const int2 a = pack3(v[0][0], v[0][1], v[0][2]), b = pack3(v[1][0], v[1][1], v[1][2]);
                const int2 c = pack3(v[2][0], v[2][1], v[2][2]), e = pack3(v[3][0], v[3][1], v[3][2]);
                *reinterpret_cast<int4*>(frow + (ptrdiff_t)X * 4) = make_int4(a.x, a.y, b.x, b.y);
                *reinterpret_cast<int4*>(frow + (ptrdiff_t)(X + 2) * 4) = make_int4(c.x, c.y, e.x, e.y);
            } else {
#pragma unroll
                for (int p = 0; p < 4; p++)
                    if (X + p >= A.sx0 && X + p < A.sx1) *reinterpret_cast<int2*>(frow + (ptrdiff_t)(X + p) * 4) = pack3(v[p][0], v[p][1], v[p][2]);
            }
        } else {
            if (yy < A.oy0 || yy >= A.oy1) continue;
            uint32_t px[4], mk = 0;
#pragma unroll
            for (int p = 0; p < 4; p++) {
                px[p] = 0;
                if (ws[dy][p] > 1e-5f) {
                    px[p] = (uint32_t)max(0, min(255, v[p][0])) | ((uint32_t)max(0, min(255, v[p][1])) << 8) | ((uint32_t)max(0, min(255, v[p][2])) << 16);
                    mk |= 0xffu << (8 * p);
                }
            }
            const size_t o = (size_t)yy * A.cw + X;
            if (X >= A.ox0 && X + 4 <= A.ox1 && (A.cw & 3) == 0) {                // 12 bytes = 3 aligned words (X and the row pitch are multiples of 4)
                const uint32_t q0 = px[0] | (px[1] << 24), q1 = (px[1] >> 8) | (px[2] << 16), q2 = (px[2] >> 16) | (px[3] << 8);
                uint32_t* o32 = reinterpret_cast<uint32_t*>(A.out + o * 3);
                o32[0] = q0; o32[1] = q1; o32[2] = q2;
                *reinterpret_cast<uint32_t*>(A.out_mask + o) = mk;
                if (A.out_peer) {                                                 // fused gather: the same 12 bytes straight into the root's mosaic (peer memory)
                    uint32_t* r32 = reinterpret_cast<uint32_t*>(A.out_peer + o * 3);
                    r32[0] = q0; r32[1] = q1; r32[2] = q2;
                }
            } else {
#pragma unroll
                for (int p = 0; p < 4; p++) {
                    if (X + p < A.ox0 || X + p >= A.ox1) continue;
                    A.out[(o + p) * 3] = (uint8_t)px[p]; A.out[(o + p) * 3 + 1] = (uint8_t)(px[p] >> 8); A.out[(o + p) * 3 + 2] = (uint8_t)(px[p] >> 16);
                    A.out_mask[o + p] = (uint8_t)(mk >> (8 * p));
                    if (A.out_peer) { A.out_peer[(o + p) * 3] = (uint8_t)px[p]; A.out_peer[(o + p) * 3 + 1] = (uint8_t)(px[p] >> 8); A.out_peer[(o + p) * 3 + 2] = (uint8_t)(px[p] >> 16); }
                }
            }
        }
    }
}

// The top level (its Laplacian is the Gaussian itself, no coarser level), one pixel per thread; tile 32 x 8.
// L0 (no bands at all): the level is the image, the result is the mosaic.
template <bool L0>
__global__ void __launch_bounds__(256)
k7_level_top(const BlendChip* __restrict__ chips, const LevelArgs A)
{
    __shared__ int list[256];
    __shared__ int wcount[8];
    const int level = A.level;
    const int tx0 = A.sx0 + blockIdx.x * 32, ty0 = A.sy0 + blockIdx.y * 8;
    const int X = tx0 + threadIdx.x, Y = ty0 + threadIdx.y;
    const bool live = X < A.sx1 && Y < A.sy1;
    int d[3] = {0, 0, 0};
    float ws = 0.0f;
    for (int base = 0; base < A.n_chips; base += 256) {
        const int total = tile_chip_list(chips, A.n_chips, base, level, tx0, ty0, tx0 + 32, ty0 + 8, list, wcount);
        for (int e = 0; e < total; e++) {
            const BlendChip& B = chips[list[e]];
            const int x = X - (B.tlx >> level), y = Y - (B.tly >> level);
            if (!live || x < B.ux0[level] || x >= B.ux1[level] || y < B.uy0[level] || y >= B.uy1[level]) continue;
            float wv; int c3[3];
            if (L0) {
                const int ix = x - B.left, iy = y - B.top;
                wv = (float)B.mask[(ptrdiff_t)iy * B.mask_step + ix] * (float)(1. / 255.);
                if (wv == 0.0f) continue;
                const uint32_t s = B.chip[(ptrdiff_t)iy * B.chip_step + ix];
                c3[0] = (int)(s & 0xffu); c3[1] = (int)((s >> 8) & 0xffu); c3[2] = (int)((s >> 16) & 0xffu);
            } else {
                const ptrdiff_t o = (ptrdiff_t)(y - B.cy0[level]) * B.cw_[level] + (x - B.cx0[level]);
                wv = B.wp[level][o];
                if (wv == 0.0f) continue;
                const uint32_t s = B.pyr[level][o];
                c3[0] = (int)(s & 0xffu); c3[1] = (int)((s >> 8) & 0xffu); c3[2] = (int)((s >> 16) & 0xffu);
            }
#pragma unroll
            for (int k = 0; k < 3; k++) d[k] = (short)(d[k] + (short)__float2int_rz((float)c3[k] * wv));
            ws += wv;
        }
        __syncthreads();
    }
    if (!live) return;
    int v[3];
#pragma unroll
    for (int k = 0; k < 3; k++) v[k] = norm_div(d[k], ws);
    if (!L0) {
        *reinterpret_cast<int2*>(A.fin + ((ptrdiff_t)(Y - A.fin_y0) * A.fin_pitch + (X - A.fin_x0)) * 4) = pack3(v[0], v[1], v[2]);
    } else if (X >= A.ox0 && X < A.ox1 && Y >= A.oy0 && Y < A.oy1) {
        const bool m = ws > 1e-5f;
        uint8_t* o = A.out + ((size_t)Y * A.cw + X) * 3;
#pragma unroll
        for (int k = 0; k < 3; k++) o[k] = m ? (uint8_t)max(0, min(255, v[k])) : 0;
        A.out_mask[(size_t)Y * A.cw + X] = m ? 255 : 0;
        if (A.out_peer) {
            uint8_t* r = A.out_peer + ((size_t)Y * A.cw + X) * 3;
#pragma unroll
            for (int k = 0; k < 3; k++) r[k] = o[k];
        }
    }
}

void free_ws(BlendWs* ws)
{
    if (!ws) return;
    cudaFree(ws->d_chips); cudaFree(ws->d_scratch);
    for (int i = 0; i <= kMaxBands; i++) cudaFree(ws->d_fin[i]);
    delete ws;
}

}  // namespace

void uavm_blend_free(uavm_canvas* cv)
{
    if (cv && cv->blend_ws) { free_ws((BlendWs*)cv->blend_ws); cv->blend_ws = nullptr; }
}

// grows a device buffer (contents are scratch)
static int ensure_cap(uavm_ctx* ctx, void** p, size_t* cap, size_t need)
{
    if (need <= *cap && *p) return UAVM_OK;
    cudaFree(*p); *p = nullptr; *cap = 0;
    UAVM_CUDA(ctx, cudaMalloc(p, need));
    *cap = need;
    return UAVM_OK;
}

// the mosaic buffers of the blend (canvas layout size), zeroed when (re)allocated
int uavm_canvas_ensure_result(uavm_ctx* ctx, uavm_canvas* cv)
{
    const int cw = cv->layout.canvas_w, ch = cv->layout.canvas_h;
    if (cv->result_w == cw && cv->result_h == ch && cv->d_result) return UAVM_OK;
    cudaFree(cv->d_result); cudaFree(cv->d_result_mask); cv->d_result = nullptr; cv->d_result_mask = nullptr;
    cv->result_w = cv->result_h = 0; cv->peer_result = nullptr;                   // a root binding refers to the old buffer
    UAVM_CUDA(ctx, cudaMalloc(&cv->d_result, (size_t)cw * ch * 3));
    UAVM_CUDA(ctx, cudaMalloc(&cv->d_result_mask, (size_t)cw * ch));
    UAVM_CUDA(ctx, cudaMemsetAsync(cv->d_result, 0, (size_t)cw * ch * 3, ctx->stream));
    UAVM_CUDA(ctx, cudaMemsetAsync(cv->d_result_mask, 0, (size_t)cw * ch, ctx->stream));
    cv->result_w = cw; cv->result_h = ch;
    return UAVM_OK;
}

extern "C" int uavm_canvas_blend(uavm_ctx* ctx, uavm_canvas* cv, int num_bands)
{
    if (!ctx || !cv || num_bands < 0 || num_bands > kMaxBands) return UAVM_EINVAL;
    if (!cv->warped) { UAVM_SET_ERR(ctx, "blend before warp"); return UAVM_EINVAL; }
    UAVM_CUDA(ctx, cudaSetDevice(ctx->device));
    { int rc = uavm_canvas_mask_plane(ctx, cv); if (rc != UAVM_OK) return rc; }     // feed masks: K6's, else the validity masks
    const int cw = cv->layout.canvas_w, ch = cv->layout.canvas_h;
    // prepare(): crop the band count, pad the canvas
    const double max_len = (double)(cw > ch ? cw : ch);
    int nb = (int)ceil(log(max_len) / log(2.0));
    if (num_bands < nb) nb = num_bands;
    if (cv->sharded && nb > 5) { UAVM_SET_ERR(ctx, "a sharded canvas supports at most 5 bands (uavm_canvas_set_rect sizes the chip regions for 5)"); return UAVM_EINVAL; }
    const IRect out = cv->sharded ? IRect{cv->rect_x0, cv->rect_y0, cv->rect_x1, cv->rect_y1} : IRect{0, 0, cw, ch};
    const CanvasPlan P = plan_canvas(cw, ch, nb, out);
    // the pixels each chip owns (K6 bounding boxes): one small read-back, the planner needs them on the host
    if (cv->seamed && !cv->own_bbox_valid) {
        cv->own_bbox.resize((size_t)cv->n * 4);
        UAVM_CUDA(ctx, cudaMemcpyAsync(cv->own_bbox.data(), cv->d_own_bbox, (size_t)cv->n * 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
        UAVM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        cv->own_bbox_valid = true;
    }
    std::vector<BlendChip> bc;
    std::vector<ChipPlan> plans;
    size_t scratch = 0;
    int max_cw[kMaxBands + 1] = {0}, max_ch[kMaxBands + 1] = {0};
    for (int k = 0; k < cv->n; k++) {
        const ChipDesc& dsc = cv->desc[k];
        if (!dsc.keep) continue;
        IRect A0{0, 0, dsc.chip_w, dsc.chip_h};
        if (cv->seamed) {
            const int32_t* b = &cv->own_bbox[(size_t)k * 4];
            A0 = (b[2] < b[0] || b[3] < b[1]) ? make_empty() : IRect{b[0], b[1], b[2] + 1, b[3] + 1};
        }
        const ChipPlan c = plan_chip(dsc.beg_x, dsc.beg_y, dsc.chip_w, dsc.chip_h, A0, P);
        if (!c.active) continue;
        BlendChip B; memset(&B, 0, sizeof(B));
        B.chip = dsc.chip; B.mask = dsc.mask; B.chip_step = dsc.chip_step; B.mask_step = dsc.mask_step; B.cw = dsc.chip_w; B.ch = dsc.chip_h;
        B.left = c.roi.left; B.top = c.roi.top; B.tlx = c.roi.tlx; B.tly = c.roi.tly;
        for (int i = 0; i <= nb; i++) {
            B.pw[i] = c.pw[i]; B.ph[i] = c.ph[i];
            B.ux0[i] = c.U[i].x0; B.uy0[i] = c.U[i].y0; B.ux1[i] = c.U[i].x1; B.uy1[i] = c.U[i].y1;
            B.cx0[i] = c.C[i].x0; B.cy0[i] = c.C[i].y0; B.cw_[i] = c.C[i].x1 - c.C[i].x0; B.ch_[i] = c.C[i].y1 - c.C[i].y0;
            if (i >= 1) {
                const size_t px = (size_t)B.cw_[i] * B.ch_[i];
                B.pyr[i] = (uint32_t*)scratch; scratch += (px * 4 + 255) & ~(size_t)255;          // offsets for now, rebased below
                B.wp[i] = (float*)scratch; scratch += (px * 4 + 255) & ~(size_t)255;
                if (B.cw_[i] > max_cw[i]) max_cw[i] = B.cw_[i];
                if (B.ch_[i] > max_ch[i]) max_ch[i] = B.ch_[i];
            }
        }
        bc.push_back(B); plans.push_back(c);
    }
    BlendWs* ws = (BlendWs*)cv->blend_ws;
    if (!ws) { ws = new BlendWs(); cv->blend_ws = ws; }
    ws->nb = nb;
    const int n_act = (int)bc.size();
    { int rc = ensure_cap(ctx, (void**)&ws->d_chips, &ws->chips_cap, (size_t)(n_act > 0 ? n_act : 1) * sizeof(BlendChip)); if (rc != UAVM_OK) return rc; }
    { int rc = ensure_cap(ctx, (void**)&ws->d_scratch, &ws->scratch_cap, scratch + 256); if (rc != UAVM_OK) return rc; }
    for (int i = 1; i <= nb; i++) {
        const size_t px = (size_t)(P.S[i].x1 - P.S[i].x0) * (P.S[i].y1 - P.S[i].y0);
        int rc = ensure_cap(ctx, (void**)&ws->d_fin[i], &ws->fin_cap[i], px * 8 + 256); if (rc != UAVM_OK) return rc;
    }
    for (auto& B : bc)
        for (int i = 1; i <= nb; i++) { B.pyr[i] = (uint32_t*)(ws->d_scratch + (size_t)B.pyr[i]); B.wp[i] = (float*)(ws->d_scratch + (size_t)B.wp[i]); }
    if (n_act > 0) UAVM_CUDA(ctx, cudaMemcpyAsync(ws->d_chips, bc.data(), (size_t)n_act * sizeof(BlendChip), cudaMemcpyHostToDevice, ctx->stream));
    { int rc = uavm_canvas_ensure_result(ctx, cv); if (rc != UAVM_OK) return rc; }
    // 1. chip pyramids, all chips per launch
    for (int i = 0; i < nb && n_act > 0; i++) {
        if (max_cw[i + 1] <= 0 || max_ch[i + 1] <= 0) continue;
        dim3 grid((max_cw[i + 1] + kPdW - 1) / kPdW, (max_ch[i + 1] + kPdH - 1) / kPdH, n_act);
        if (grid.y > 65535) { UAVM_SET_ERR(ctx, "blend: chip pyramid level %d too tall", i + 1); return UAVM_EINVAL; }
        if (i == 0) k7_pyrdown<true><<<grid, dim3(kPdW, 4), 0, ctx->stream>>>(ws->d_chips, 0);
        else k7_pyrdown<false><<<grid, dim3(kPdW, 4), 0, ctx->stream>>>(ws->d_chips, i);
        UAVM_CHECK_LAUNCH(ctx);
    }
    // 2. canvas levels, top down
    for (int i = nb; i >= 0; i--) {
        LevelArgs A; memset(&A, 0, sizeof(A));
        A.level = i; A.n_chips = n_act;
        A.sx0 = P.S[i].x0; A.sy0 = P.S[i].y0; A.sx1 = P.S[i].x1; A.sy1 = P.S[i].y1;
        if (A.sx1 <= A.sx0 || A.sy1 <= A.sy0) continue;
        if (i >= 1) { A.fin = ws->d_fin[i]; A.fin_x0 = P.S[i].x0; A.fin_y0 = P.S[i].y0; A.fin_pitch = P.S[i].x1 - P.S[i].x0; }
        if (i < nb) { A.nxt = ws->d_fin[i + 1]; A.nxt_x0 = P.S[i + 1].x0; A.nxt_y0 = P.S[i + 1].y0; A.nxt_pitch = P.S[i + 1].x1 - P.S[i + 1].x0; A.nxt_rows = P.S[i + 1].y1 - P.S[i + 1].y0; A.nxt_w = P.lw[i + 1]; A.nxt_h = P.lh[i + 1]; }
        if (i == 0) {
            A.out = cv->d_result; A.out_mask = cv->d_result_mask; A.cw = cw; A.ch = ch;
            A.out_peer = cv->peer_result;
            A.ox0 = out.x0; A.oy0 = out.y0; A.ox1 = out.x1 < cw ? out.x1 : cw; A.oy1 = out.y1 < ch ? out.y1 : ch;
        }
        if (i == nb) {
            dim3 grid((A.sx1 - A.sx0 + 31) / 32, (A.sy1 - A.sy0 + 7) / 8);
            if (grid.y > 65535) { UAVM_SET_ERR(ctx, "blend: canvas level %d too tall", i); return UAVM_EINVAL; }
            if (i == 0) k7_level_top<true><<<grid, dim3(32, 8), 0, ctx->stream>>>(ws->d_chips, A);
            else k7_level_top<false><<<grid, dim3(32, 8), 0, ctx->stream>>>(ws->d_chips, A);
        } else {
            dim3 grid((A.sx1 - (A.sx0 & ~3) + 127) / 128, (A.sy1 - A.sy0 + 15) / 16);
            if (grid.y > 65535) { UAVM_SET_ERR(ctx, "blend: canvas level %d too tall", i); return UAVM_EINVAL; }
            static const int minb = getenv("UAVM_K7_MINB") ? atoi(getenv("UAVM_K7_MINB")) : 4;        // A/B knob: resident CTAs per SM the compiler targets
            if (minb == 3) { if (i == 0) k7_level<true, 3><<<grid, dim3(32, 8), 0, ctx->stream>>>(ws->d_chips, A); else k7_level<false, 3><<<grid, dim3(32, 8), 0, ctx->stream>>>(ws->d_chips, A); }
            else if (minb == 4) { if (i == 0) k7_level<true, 4><<<grid, dim3(32, 8), 0, ctx->stream>>>(ws->d_chips, A); else k7_level<false, 4><<<grid, dim3(32, 8), 0, ctx->stream>>>(ws->d_chips, A); }
            else { if (i == 0) k7_level<true, 2><<<grid, dim3(32, 8), 0, ctx->stream>>>(ws->d_chips, A); else k7_level<false, 2><<<grid, dim3(32, 8), 0, ctx->stream>>>(ws->d_chips, A); }
        }
        UAVM_CHECK_LAUNCH(ctx);
    }
    cv->blended = true;
    cv->blend_bound_root = cv->bound_root;
    return UAVM_OK;
}

extern "C" int uavm_canvas_get_result(uavm_ctx* ctx, uavm_canvas* cv, uint8_t* bgr, int step, uint8_t* mask, int mask_step)
{
    if (!ctx || !cv) return UAVM_EINVAL;
    if (!cv->blended || !cv->d_result) { UAVM_SET_ERR(ctx, "get_result before blend"); return UAVM_EINVAL; }
    const int cw = cv->result_w, ch = cv->result_h;
    if (bgr) {
        if (step < cw * 3) return UAVM_EINVAL;
        UAVM_CUDA(ctx, cudaMemcpy2DAsync(bgr, (size_t)step, cv->d_result, (size_t)cw * 3, (size_t)cw * 3, ch, cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (mask) {
        if (mask_step < cw) return UAVM_EINVAL;
        UAVM_CUDA(ctx, cudaMemcpy2DAsync(mask, (size_t)mask_step, cv->d_result_mask, (size_t)cw, (size_t)cw, ch, cudaMemcpyDeviceToHost, ctx->stream));
    }
    UAVM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return UAVM_OK;
}

extern "C" int uavm_canvas_result_size(uavm_canvas* cv, int* w, int* h)
{
    if (!cv || !w || !h) return UAVM_EINVAL;
    *w = cv->result_w; *h = cv->result_h;
    return UAVM_OK;
}

// rows [y0, y1) of the result (dense, canvas_w * 3 bytes per row) into a caller buffer; is_device != 0: device
// memory (stream-ordered device-to-device copy, e.g. into the send buffer of an NCCL gather of canvas bands)
extern "C" int uavm_canvas_copy_result_rows(uavm_ctx* ctx, uavm_canvas* cv, int y0, int y1, uint8_t* dst, int is_device)
{
    if (!ctx || !cv || !dst || y0 < 0 || y1 > cv->result_h || y0 > y1) return UAVM_EINVAL;
    if (!cv->blended || !cv->d_result) { UAVM_SET_ERR(ctx, "copy_result_rows before blend"); return UAVM_EINVAL; }
    const size_t row = (size_t)cv->result_w * 3;
    if (y1 > y0)
        UAVM_CUDA(ctx, cudaMemcpyAsync(dst, cv->d_result + (size_t)y0 * row, (size_t)(y1 - y0) * row,
                                       is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, ctx->stream));
    if (!is_device) UAVM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return UAVM_OK;
}

// rectangle [x0, x1) x [y0, y1) of the result, dense ((x1 - x0) * 3 bytes per row): the piece a rank of a 2-D sharded canvas
// contributes to the gather
extern "C" int uavm_canvas_copy_result_rect(uavm_ctx* ctx, uavm_canvas* cv, int x0, int y0, int x1, int y1, uint8_t* dst, int is_device)
{
    if (!ctx || !cv || !dst || x0 < 0 || y0 < 0 || x1 > cv->result_w || y1 > cv->result_h || x0 > x1 || y0 > y1) return UAVM_EINVAL;
    if (!cv->blended || !cv->d_result) { UAVM_SET_ERR(ctx, "copy_result_rect before blend"); return UAVM_EINVAL; }
    if (x1 > x0 && y1 > y0)
        UAVM_CUDA(ctx, cudaMemcpy2DAsync(dst, (size_t)(x1 - x0) * 3, cv->d_result + ((size_t)y0 * cv->result_w + x0) * 3, (size_t)cv->result_w * 3,
                                         (size_t)(x1 - x0) * 3, (size_t)(y1 - y0), is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, ctx->stream));
    if (!is_device) UAVM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return UAVM_OK;
}
