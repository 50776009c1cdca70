// blend.cu — K7: multi-band blend.  Replaces detail::MultiBandBlender prepare / feed / blend and the final
// convertTo(CV_8U) as driven by LaplacianPyramidBlending (M/MosaicImage.cpp:2296-2299, :2476-2486).
//
// The blender itself is OpenCV code (third party, not under /root/reference); the arithmetic restated here is
// the one documented in oracle/oracle_blend.c: int16 Laplacian pyramids ([1 4 6 4 1] pyrDown with
// (sum+128)>>8, pyrUp with (sum+32)>>6, reflect-101 / replicate borders), f32 Gaussian weight pyramids,
// dst += short(src * w), wsum += w, dst = short(dst / (wsum + 1e-5)), collapse with saturating adds.
//
// HBM layout: the canvas pyramid (int16 x4 pixels = B, G, R, pad + f32 weights per level, canvas padded to a multiple
// of 2^bands) stays resident; every fed chip gets a scratch pyramid of its padded ROI that is reused for the
// next chip.  Images are fed in index order (the f32 weight sums are order dependent), pixels in parallel.
#include <math.h>
#include <string.h>
#include "canvas.h"

namespace {

constexpr int kMaxBands = 8;

struct BlendWs {
    int nb = 0, W = 0, H = 0, Y0 = 0, Y1 = 0;   // Y0..Y1: canvas rows held by this workspace (band + halo)
    int lw[kMaxBands + 1], lh[kMaxBands + 1];
    short* dlap[kMaxBands + 1] = {nullptr};
    float* dw[kMaxBands + 1] = {nullptr};
    short* pyr[kMaxBands + 1] = {nullptr};     // scratch pyramid of the chip being fed (capacity of the largest ROI)
    float* wp[kMaxBands + 1] = {nullptr};
    size_t roi_cap = 0;
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
__device__ __forceinline__ short sat16(int v) { return (short)max(-32768, min(32767, v)); }

// Pyramid pixels are 4 x int16 (B, G, R, 0): one aligned 8-byte load / store per pixel instead of three 2-byte ones
// (a 6-byte interleaved pixel straddles words; the kernels were LSU-instruction bound with it).
__device__ __forceinline__ void ld3(const short* base, size_t px, int& a, int& b, int& c)
{
    const int2 v = *reinterpret_cast<const int2*>(base + px * 4);
    a = (short)(v.x & 0xffff); b = v.x >> 16; c = (short)(v.y & 0xffff);
}
__device__ __forceinline__ void st3(short* base, size_t px, int a, int b, int c)
{
    int2 v; v.x = (a & 0xffff) | (b << 16); v.y = c & 0xffff;
    *reinterpret_cast<int2*>(base + px * 4) = v;
}

// level 0 of the fed image: copyMakeBorder(REFLECT) of the chip (u8 -> s16) and mask/255 with a zero border.
// Four pixels per thread; groups that lie inside the chip with 4-pixel alignment on both sides move as vectors.
__global__ void __launch_bounds__(256)
k7_feed_level0(const uint32_t* __restrict__ chip, int chip_step /* words */, const uint8_t* __restrict__ mask, int mask_step,
               int cw, int ch, int left, int top, int width, int height, short* __restrict__ pyr0, float* __restrict__ wp0)
{
    const int x0 = 4 * (blockIdx.x * blockDim.x + threadIdx.x), y = blockIdx.y;
    if (x0 >= width || y >= height) return;
    const int iy = y - top, sy = reflect_edge(iy, ch);
    const int ix0 = x0 - left;
    const size_t o = (size_t)y * width + x0;
    if (ix0 >= 0 && ix0 + 4 <= cw && x0 + 4 <= width && (width & 3) == 0) {
        const uint32_t* crow = chip + (size_t)sy * chip_step + ix0;                                    // 4 BGRA pixels
        uint32_t sv[4];
        if ((ix0 & 3) == 0) { const uint4 s = *reinterpret_cast<const uint4*>(crow); sv[0] = s.x; sv[1] = s.y; sv[2] = s.z; sv[3] = s.w; }
        else { sv[0] = crow[0]; sv[1] = crow[1]; sv[2] = crow[2]; sv[3] = crow[3]; }
        int4 a, b;
        a.x = (int)((sv[0] & 0xffu) | ((sv[0] & 0xff00u) << 8)); a.y = (int)((sv[0] >> 16) & 0xffu);
        a.z = (int)((sv[1] & 0xffu) | ((sv[1] & 0xff00u) << 8)); a.w = (int)((sv[1] >> 16) & 0xffu);
        b.x = (int)((sv[2] & 0xffu) | ((sv[2] & 0xff00u) << 8)); b.y = (int)((sv[2] >> 16) & 0xffu);
        b.z = (int)((sv[3] & 0xffu) | ((sv[3] & 0xff00u) << 8)); b.w = (int)((sv[3] >> 16) & 0xffu);
        *reinterpret_cast<int4*>(pyr0 + o * 4) = a; *reinterpret_cast<int4*>(pyr0 + (o + 2) * 4) = b;
        float4 wv = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (iy >= 0 && iy < ch) {
            // 4 mask bytes at an arbitrary byte offset: two aligned words (rows are padded to align4, + 256 B slack) and a funnel shift
            const uint32_t* mrow = reinterpret_cast<const uint32_t*>(mask + (size_t)iy * mask_step + (ix0 & ~3));
            const uint32_t m = (ix0 & 3) ? __funnelshift_r(mrow[0], mrow[1], 8 * (ix0 & 3)) : mrow[0];
            wv.x = (float)(m & 0xffu) * (float)(1. / 255.); wv.y = (float)((m >> 8) & 0xffu) * (float)(1. / 255.);
            wv.z = (float)((m >> 16) & 0xffu) * (float)(1. / 255.); wv.w = (float)(m >> 24) * (float)(1. / 255.);
        }
        *reinterpret_cast<float4*>(wp0 + o) = wv;
        return;
    }
    for (int i = 0; i < 4 && x0 + i < width; i++) {
        const int ix = ix0 + i;
        const int sx = reflect_edge(ix, cw);
        const uint32_t s = chip[(size_t)sy * chip_step + sx];                   // BGRA
        st3(pyr0, o + i, (int)(s & 0xffu), (int)((s >> 8) & 0xffu), (int)((s >> 16) & 0xffu));
        float wv = 0.0f;
        if (ix >= 0 && ix < cw && iy >= 0 && iy < ch) wv = (float)mask[(size_t)iy * mask_step + ix] * (float)(1. / 255.);
        wp0[o + i] = wv;
    }
}

// pyrDown of the int16 x3 image and of the f32 weight map, one output pixel per thread
__global__ void __launch_bounds__(256)
k7_pyrdown(const short* __restrict__ src, const float* __restrict__ wsrc, int w, int h,
           short* __restrict__ dst, float* __restrict__ wdst, int dw, int dh)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= dw || y >= dh) return;
    int xs[5], ys[5];
#pragma unroll
    for (int k = 0; k < 5; k++) { xs[k] = reflect101(2 * x - 2 + k, w); ys[k] = reflect101(2 * y - 2 + k, h); }
    int acc[3] = {0, 0, 0};
    float frow[5];
    const int kw[5] = {1, 4, 6, 4, 1};
#pragma unroll
    for (int r = 0; r < 5; r++) {
        const size_t rbase = (size_t)ys[r] * w;
        int ra[3] = {0, 0, 0};
#pragma unroll
        for (int k = 0; k < 5; k++) {
            int p0, p1, p2;
            ld3(src, rbase + xs[k], p0, p1, p2);
            ra[0] += kw[k] * p0; ra[1] += kw[k] * p1; ra[2] += kw[k] * p2;
        }
        acc[0] += kw[r] * ra[0]; acc[1] += kw[r] * ra[1]; acc[2] += kw[r] * ra[2];
        if (wsrc) {
            const float* wr = wsrc + rbase;
            const float s0 = wr[xs[0]], s1 = wr[xs[1]], s2 = wr[xs[2]], s3 = wr[xs[3]], s4 = wr[xs[4]];
            frow[r] = s2 * 6.0f + (s1 + s3) * 4.0f + s0 + s4;           // row pass, oracle order
        }
    }
    st3(dst, (size_t)y * dw + x, sat16((acc[0] + 128) >> 8), sat16((acc[1] + 128) >> 8), sat16((acc[2] + 128) >> 8));
    if (wsrc) {
        const float v = frow[2] * 6.0f + (frow[1] + frow[3]) * 4.0f + frow[0] + frow[4];   // column pass
        wdst[(size_t)y * dw + x] = v * (1.0f / 256.0f);
    }
}

// Tiled pyrDown: a CTA produces a 64 x 8 output tile from a 131 x 19 source window staged in shared memory with coalesced
// loads (every source line is requested once instead of ~5 times through L1), split into even / odd columns so that the
// stride-2 taps become unit-stride, conflict-free LDS.  A thread computes two vertically adjacent outputs: 7 window rows,
// each filtered horizontally once (integer sums are order free; the f32 passes keep the oracle's expression order).
constexpr int kPdW = 64, kPdH = 8, kPdCols = 2 * kPdW + 4 /* 131 used, even/odd halves of 66 */, kPdRows = 2 * kPdH + 3;
__global__ void __launch_bounds__(256)
k7_pyrdown_tiled(const short* __restrict__ src, const float* __restrict__ wsrc, int w, int h,
                 short* __restrict__ dst, float* __restrict__ wdst, int dw, int dh)
{
    __shared__ int2 sE[kPdRows][kPdCols / 2], sO[kPdRows][kPdCols / 2];      // pixel = 4 x int16 = int2
    __shared__ float wE[kPdRows][kPdCols / 2], wO[kPdRows][kPdCols / 2];
    const int tid = threadIdx.y * kPdW + threadIdx.x;
    const int ox0 = blockIdx.x * kPdW, oy0 = blockIdx.y * kPdH;
    const int sx0 = 2 * ox0 - 2, sy0 = 2 * oy0 - 2;                           // window origin (even column)
    const bool even_w = (w & 1) == 0;
    for (int e = tid; e < kPdRows * (kPdCols / 2); e += 256) {               // one (even, odd) column pair per step
        const int r = e / (kPdCols / 2), j = e - r * (kPdCols / 2);
        const size_t rb = (size_t)reflect101(sy0 + r, h) * w;
        const int gx = sx0 + 2 * j;
        if (even_w && gx >= 0 && gx + 1 < w) {                                // aligned 16-byte pixel pair + 8-byte weight pair
            const int4 v = *reinterpret_cast<const int4*>(src + (rb + gx) * 4);
            sE[r][j] = make_int2(v.x, v.y); sO[r][j] = make_int2(v.z, v.w);
            float2 f = make_float2(0.0f, 0.0f);
            if (wsrc) f = *reinterpret_cast<const float2*>(wsrc + rb + gx);
            wE[r][j] = f.x; wO[r][j] = f.y;
        } else {                                                              // border columns: reflect-101 per pixel
            const size_t g0 = rb + reflect101(gx, w), g1 = rb + reflect101(gx + 1, w);
            sE[r][j] = *reinterpret_cast<const int2*>(src + g0 * 4); sO[r][j] = *reinterpret_cast<const int2*>(src + g1 * 4);
            wE[r][j] = wsrc ? wsrc[g0] : 0.0f; wO[r][j] = wsrc ? wsrc[g1] : 0.0f;
        }
    }
    __syncthreads();
    const int x = ox0 + threadIdx.x, ty = threadIdx.y;                        // outputs (x, oy0 + 2 ty) and (x, oy0 + 2 ty + 1)
    if (x >= dw) return;
    const int i = threadIdx.x;                                                // window column 2 i .. 2 i + 4 = E[i], O[i], E[i+1], O[i+1], E[i+2]
    int acc[2][3] = {{0, 0, 0}, {0, 0, 0}};
    float F[7];
    const int kw[5] = {1, 4, 6, 4, 1};
#pragma unroll
    for (int r = 0; r < 7; r++) {
        const int wr = 4 * ty + r;                                            // window row
        const int2 p0 = sE[wr][i], p1 = sO[wr][i], p2 = sE[wr][i + 1], p3 = sO[wr][i + 1], p4 = sE[wr][i + 2];
        int hs[3];
        hs[0] = (short)(p0.x & 0xffff) + (short)(p4.x & 0xffff) + 4 * ((short)(p1.x & 0xffff) + (short)(p3.x & 0xffff)) + 6 * (short)(p2.x & 0xffff);
        hs[1] = (p0.x >> 16) + (p4.x >> 16) + 4 * ((p1.x >> 16) + (p3.x >> 16)) + 6 * (p2.x >> 16);
        hs[2] = (short)(p0.y & 0xffff) + (short)(p4.y & 0xffff) + 4 * ((short)(p1.y & 0xffff) + (short)(p3.y & 0xffff)) + 6 * (short)(p2.y & 0xffff);
#pragma unroll
        for (int c = 0; c < 3; c++) {
            if (r < 5) acc[0][c] += kw[r] * hs[c];
            if (r >= 2) acc[1][c] += kw[r - 2] * hs[c];
        }
        if (wsrc) F[r] = wE[wr][i + 1] * 6.0f + (wO[wr][i] + wO[wr][i + 1]) * 4.0f + wE[wr][i] + wE[wr][i + 2];     // row pass, oracle order
    }
#pragma unroll
    for (int o = 0; o < 2; o++) {
        const int y = oy0 + 2 * ty + o;
        if (y >= dh) break;
        st3(dst, (size_t)y * dw + x, sat16((acc[o][0] + 128) >> 8), sat16((acc[o][1] + 128) >> 8), sat16((acc[o][2] + 128) >> 8));
        if (wsrc) {
            const float v = F[2 + 2 * o] * 6.0f + (F[1 + 2 * o] + F[3 + 2 * o]) * 4.0f + F[0 + 2 * o] + F[4 + 2 * o];  // column pass
            wdst[(size_t)y * dw + x] = v * (1.0f / 256.0f);
        }
    }
}

// value of pyrUp(lo) at (x, y) of the 2x larger level; lo is lw x lh, 3 channels
__device__ __forceinline__ void pyrup_at(const short* __restrict__ lo, int lw, int lh, int x, int y, int out[3])
{
    const int cx = x >> 1, cy = y >> 1;
    const int xm = (cx == 0) ? (lw > 1 ? 1 : 0) : cx - 1, xp = (cx == lw - 1) ? lw - 1 : cx + 1;
    const int ym = (cy == 0) ? (lh > 1 ? 1 : 0) : cy - 1, yp = (cy == lh - 1) ? lh - 1 : cy + 1;
    // per axis: even sample s[i-1] + 6 s[i] + s[i+1], odd sample 4 (s[i] + s[i+1])
    const int wx0 = (x & 1) ? 0 : 1, wx1 = (x & 1) ? 4 : 6, wx2 = (x & 1) ? 4 : 1;
    const int wy0 = (y & 1) ? 0 : 1, wy1 = (y & 1) ? 4 : 6, wy2 = (y & 1) ? 4 : 1;
    const int rows[3] = {ym, cy, yp}, wy[3] = {wy0, wy1, wy2};
    int acc[3] = {0, 0, 0};
#pragma unroll
    for (int r = 0; r < 3; r++) {
        if (wy[r] == 0) continue;
        const size_t rbase = (size_t)rows[r] * lw;
        int a[3], b[3], c[3];
        if (wx0) ld3(lo, rbase + xm, a[0], a[1], a[2]); else { a[0] = a[1] = a[2] = 0; }
        ld3(lo, rbase + cx, b[0], b[1], b[2]);
        ld3(lo, rbase + xp, c[0], c[1], c[2]);
#pragma unroll
        for (int k = 0; k < 3; k++) acc[k] += wy[r] * (wx0 * a[k] + wx1 * b[k] + wx2 * c[k]);
    }
#pragma unroll
    for (int k = 0; k < 3; k++) out[k] = sat16((acc[k] + 32) >> 6);
}

// Laplacian level (pyr[i] - pyrUp(pyr[i+1]), saturating; the top level is the Gaussian itself) times the
// weight, accumulated into the canvas pyramid: dst += short(lap * w), wsum += w
__global__ void __launch_bounds__(256)
k7_lap_accumulate(const short* __restrict__ cur, const short* __restrict__ next, const float* __restrict__ wcur,
                  int w, int h, int nw, int nh, short* __restrict__ dlap, float* __restrict__ dwsum, int dst_w, int x_tl, int y_tl)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w || y >= h) return;
    // Seam masks give every canvas pixel one owner, so most of a fed chip has weight exactly 0 (about 2/3 of level 0
    // for a 3x-covered strip): dst + short(lap * 0) == dst and wsum + 0 == wsum, the read-modify-write is skipped.
    if (wcur[(size_t)y * w + x] == 0.0f) return;
    int lap[3];
    ld3(cur, (size_t)y * w + x, lap[0], lap[1], lap[2]);
    if (next) {
        int up[3];
        pyrup_at(next, nw, nh, x, y, up);
#pragma unroll
        for (int k = 0; k < 3; k++) lap[k] = sat16(lap[k] - up[k]);
    }
    const float wv = wcur[(size_t)y * w + x];
    const size_t di = (size_t)(y + y_tl) * dst_w + (x + x_tl);
    int d[3];
    ld3(dlap, di, d[0], d[1], d[2]);
#pragma unroll
    for (int k = 0; k < 3; k++) d[k] = (short)(d[k] + (short)__float2int_rz((float)lap[k] * wv));
    st3(dlap, di, d[0], d[1], d[2]);
    dwsum[di] += wv;
}

// blend(): dst = short(dst / (wsum + 1e-5)) per level, then restoreImageFromLaplacePyr: hi = sat(pyrUp(lo) + hi) from the
// top down.  The normalisation of level i is fused into the collapse step that consumes it (lo is already final), and
// at level 0 the step writes the u8 mosaic directly: crop, zero where wsum <= 1e-5, convertTo(CV_8U) (saturate).
__device__ __forceinline__ void normalized3(const short* lap, const float* wsum, size_t i, int d[3])
{
    const float wv = wsum[i] + 1e-5f;
    ld3(lap, i, d[0], d[1], d[2]);
#pragma unroll
    for (int k = 0; k < 3; k++) d[k] = (short)__float2int_rz((float)d[k] / wv);
}

// Same for the levels that have a coarser level below them, one 2 x 2 quad of pixels per thread: the four pixels share
// the 3 x 3 neighbourhood of the coarser level (9 loads instead of 9 + 6 + 6 + 4) and pyrUp separates into three horizontal
// sums per parity and a vertical combination; cur / dst / weights move as aligned 16- and 8-byte vectors.  Level sizes and
// the paste offsets (x_tl, y_tl) are even below the top level (ROIs are aligned to 2^bands), so quads tile exactly.
__global__ void __launch_bounds__(256)
k7_lap_accumulate_quad(const short* __restrict__ cur, const short* __restrict__ next, const float* __restrict__ wcur,
                       int w, int h, int nw, int nh, short* __restrict__ dlap, float* __restrict__ dwsum, int dst_w, int x_tl, int y_tl)
{
    const int cx = blockIdx.x * blockDim.x + threadIdx.x, cy = blockIdx.y;          // quad index = coordinates in the coarser level
    const int x = 2 * cx, y = 2 * cy;
    if (x >= w || y >= h) return;
    const float2 w0 = *reinterpret_cast<const float2*>(wcur + (size_t)y * w + x);
    const float2 w1 = *reinterpret_cast<const float2*>(wcur + (size_t)(y + 1) * w + x);
    if (w0.x == 0.0f && w0.y == 0.0f && w1.x == 0.0f && w1.y == 0.0f) return;       // dst + short(lap * 0) == dst, wsum + 0 == wsum
    // pyrUp border rules (see pyrup_at): reflect-101 at the near edge, replicate at the far edge
    const int xm = (cx == 0) ? (nw > 1 ? 1 : 0) : cx - 1, xp = (cx == nw - 1) ? nw - 1 : cx + 1;
    const int ym = (cy == 0) ? (nh > 1 ? 1 : 0) : cy - 1, yp = (cy == nh - 1) ? nh - 1 : cy + 1;
    const int rows[3] = {ym, cy, yp};
    int he[3][3], ho[3][3];                        // [row][channel]: even-x sum a + 6 b + c, odd-x sum 4 (b + c)
#pragma unroll
    for (int r = 0; r < 3; r++) {
        const size_t rb = (size_t)rows[r] * nw;
        int a[3], b[3], c[3];
        ld3(next, rb + xm, a[0], a[1], a[2]); ld3(next, rb + cx, b[0], b[1], b[2]); ld3(next, rb + xp, c[0], c[1], c[2]);
#pragma unroll
        for (int k = 0; k < 3; k++) { he[r][k] = a[k] + 6 * b[k] + c[k]; ho[r][k] = 4 * (b[k] + c[k]); }
    }
    const float wq[2][2] = {{w0.x, w0.y}, {w1.x, w1.y}};
#pragma unroll
    for (int dy = 0; dy < 2; dy++) {
        const int4 c4 = *reinterpret_cast<const int4*>(cur + ((size_t)(y + dy) * w + x) * 4);
        const size_t di = (size_t)(y + dy + y_tl) * dst_w + (x + x_tl);
        int4 d4 = *reinterpret_cast<const int4*>(dlap + di * 4);
        float2 ws = *reinterpret_cast<const float2*>(dwsum + di);
        const int cw2[2][2] = {{c4.x, c4.y}, {c4.z, c4.w}};
        int dw2[2][2] = {{d4.x, d4.y}, {d4.z, d4.w}};
#pragma unroll
        for (int dx = 0; dx < 2; dx++) {
            const int cc[3] = {(short)(cw2[dx][0] & 0xffff), cw2[dx][0] >> 16, (short)(cw2[dx][1] & 0xffff)};
            int dd[3] = {(short)(dw2[dx][0] & 0xffff), dw2[dx][0] >> 16, (short)(dw2[dx][1] & 0xffff)};
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const int h0 = dx ? ho[0][k] : he[0][k], h1 = dx ? ho[1][k] : he[1][k], h2 = dx ? ho[2][k] : he[2][k];
                const int acc = dy ? 4 * (h1 + h2) : h0 + 6 * h1 + h2;
                const int lap = sat16(cc[k] - sat16((acc + 32) >> 6));
                dd[k] = (short)(dd[k] + (short)__float2int_rz((float)lap * wq[dy][dx]));
            }
            dw2[dx][0] = (dd[0] & 0xffff) | (dd[1] << 16); dw2[dx][1] = dd[2] & 0xffff;
        }
        d4.x = dw2[0][0]; d4.y = dw2[0][1]; d4.z = dw2[1][0]; d4.w = dw2[1][1];
        ws.x += wq[dy][0]; ws.y += wq[dy][1];
        *reinterpret_cast<int4*>(dlap + di * 4) = d4;
        *reinterpret_cast<float2*>(dwsum + di) = ws;
    }
}

__global__ void __launch_bounds__(256)
k7_normalize(short* __restrict__ dlap, const float* __restrict__ dwsum, size_t n)        // top level only
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int d[3];
    normalized3(dlap, dwsum, i, d);
    st3(dlap, i, d[0], d[1], d[2]);
}

// pyrUp of the coarser level for the 2 x 2 quad (cx, cy): up[dy][dx][channel]  (same separation as k7_lap_accumulate_quad)
__device__ __forceinline__ void pyrup_quad(const short* __restrict__ lo, int lw, int lh, int cx, int cy, int up[2][2][3])
{
    const int xm = (cx == 0) ? (lw > 1 ? 1 : 0) : cx - 1, xp = (cx == lw - 1) ? lw - 1 : cx + 1;
    const int ym = (cy == 0) ? (lh > 1 ? 1 : 0) : cy - 1, yp = (cy == lh - 1) ? lh - 1 : cy + 1;
    const int rows[3] = {ym, cy, yp};
    int he[3][3], ho[3][3];
#pragma unroll
    for (int r = 0; r < 3; r++) {
        const size_t rb = (size_t)rows[r] * lw;
        int a[3], b[3], c[3];
        ld3(lo, rb + xm, a[0], a[1], a[2]); ld3(lo, rb + cx, b[0], b[1], b[2]); ld3(lo, rb + xp, c[0], c[1], c[2]);
#pragma unroll
        for (int k = 0; k < 3; k++) { he[r][k] = a[k] + 6 * b[k] + c[k]; ho[r][k] = 4 * (b[k] + c[k]); }
    }
#pragma unroll
    for (int dx = 0; dx < 2; dx++)
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const int h0 = dx ? ho[0][k] : he[0][k], h1 = dx ? ho[1][k] : he[1][k], h2 = dx ? ho[2][k] : he[2][k];
            up[0][dx][k] = sat16((h0 + 6 * h1 + h2 + 32) >> 6);
            up[1][dx][k] = sat16((4 * (h1 + h2) + 32) >> 6);
        }
}

// collapse step for even level sizes: one 2 x 2 quad per thread
__global__ void __launch_bounds__(256)
k7_collapse_quad(const short* __restrict__ lo, int lw, int lh, short* __restrict__ hi, const float* __restrict__ hi_wsum, int w, int h)
{
    const int cx = blockIdx.x * blockDim.x + threadIdx.x, cy = blockIdx.y;
    const int x = 2 * cx, y = 2 * cy;
    if (x >= w || y >= h) return;
    int up[2][2][3];
    pyrup_quad(lo, lw, lh, cx, cy, up);
#pragma unroll
    for (int dy = 0; dy < 2; dy++)
#pragma unroll
        for (int dx = 0; dx < 2; dx++) {
            int d[3];
            const size_t i = (size_t)(y + dy) * w + x + dx;
            normalized3(hi, hi_wsum, i, d);
            st3(hi, i, sat16(up[dy][dx][0] + d[0]), sat16(up[dy][dx][1] + d[1]), sat16(up[dy][dx][2] + d[2]));
        }
}

__global__ void __launch_bounds__(256)
k7_collapse(const short* __restrict__ lo, int lw, int lh, short* __restrict__ hi, const float* __restrict__ hi_wsum, int w, int h)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w || y >= h) return;
    int up[3], d[3];
    pyrup_at(lo, lw, lh, x, y, up);
    normalized3(hi, hi_wsum, (size_t)y * w + x, d);
    st3(hi, (size_t)y * w + x, sat16(up[0] + d[0]), sat16(up[1] + d[1]), sat16(up[2] + d[2]));
}

// last step (level 0) fused with the output: rows [oy0, oy1) x columns [0, cw) of the padded level go to the mosaic.
// A thread owns a 4 x 2 block (two quads): per row the 12 output bytes are three aligned words when the mosaic row pitch
// allows it.  oy0 is even (band edges are multiples of 32).
__global__ void __launch_bounds__(256)
k7_collapse_output(const short* __restrict__ lo, int lw, int lh, const short* __restrict__ hi, const float* __restrict__ hi_wsum, int w, int h,
                   int oy0, int oy1, int cw, uint8_t* __restrict__ out, uint8_t* __restrict__ out_mask)
{
    const int x0 = 4 * (blockIdx.x * blockDim.x + threadIdx.x), y0 = oy0 + 2 * blockIdx.y;
    if (x0 >= cw || y0 >= oy1) return;
    uint32_t px[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};   // B | G << 8 | R << 16
    uint32_t mk[2] = {0, 0};
#pragma unroll
    for (int q = 0; q < 2; q++) {
        const int x = x0 + 2 * q;
        if (x >= w) break;                              // w is even and >= cw
        float ws[2][2];
        bool any = false;
#pragma unroll
        for (int dy = 0; dy < 2; dy++) {
            const float2 f = *reinterpret_cast<const float2*>(hi_wsum + (size_t)(y0 + dy) * w + x);
            ws[dy][0] = f.x; ws[dy][1] = f.y; any = any || f.x > 1e-5f || f.y > 1e-5f;
        }
        if (!any) continue;
        int up[2][2][3];
        pyrup_quad(lo, lw, lh, x >> 1, y0 >> 1, up);
#pragma unroll
        for (int dy = 0; dy < 2; dy++)
#pragma unroll
            for (int dx = 0; dx < 2; dx++) {
                if (!(ws[dy][dx] > 1e-5f)) continue;
                int d[3];
                normalized3(hi, hi_wsum, (size_t)(y0 + dy) * w + x + dx, d);
                const int b = max(0, min(255, (int)sat16(up[dy][dx][0] + d[0]))), g = max(0, min(255, (int)sat16(up[dy][dx][1] + d[1]))),
                          r = max(0, min(255, (int)sat16(up[dy][dx][2] + d[2])));
                px[dy][2 * q + dx] = (uint32_t)b | ((uint32_t)g << 8) | ((uint32_t)r << 16);
                mk[dy] |= 0xffu << (8 * (2 * q + dx));
            }
    }
#pragma unroll
    for (int dy = 0; dy < 2; dy++) {
        if (y0 + dy >= oy1) break;
        const size_t o = (size_t)(2 * blockIdx.y + dy) * cw + x0;
        if (x0 + 4 <= cw && (cw & 3) == 0) {          // 12 bytes = 3 words, rows are 4-byte aligned
            uint32_t* o32 = reinterpret_cast<uint32_t*>(out + o * 3);
            o32[0] = px[dy][0] | (px[dy][1] << 24); o32[1] = (px[dy][1] >> 8) | (px[dy][2] << 16); o32[2] = (px[dy][2] >> 16) | (px[dy][3] << 8);
            *reinterpret_cast<uint32_t*>(out_mask + o) = mk[dy];
        } else {
            for (int i = 0; i < 4 && x0 + i < cw; i++) {
                out[(o + i) * 3] = (uint8_t)px[dy][i]; out[(o + i) * 3 + 1] = (uint8_t)(px[dy][i] >> 8); out[(o + i) * 3 + 2] = (uint8_t)(px[dy][i] >> 16);
                out_mask[o + i] = (uint8_t)(mk[dy] >> (8 * i));
            }
        }
    }
}

// no bands: the canvas level 0 is the (normalised) image itself
__global__ void __launch_bounds__(256)
k7_output(const short* __restrict__ lap0, const float* __restrict__ w0, int W, int cw, int ch,
          uint8_t* __restrict__ out, uint8_t* __restrict__ out_mask)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= cw || y >= ch) return;
    const bool m = w0[(size_t)y * W + x] > 1e-5f;
    int s[3];
    ld3(lap0, (size_t)y * W + x, s[0], s[1], s[2]);
    uint8_t* d = out + ((size_t)y * cw + x) * 3;
#pragma unroll
    for (int k = 0; k < 3; k++) d[k] = m ? (uint8_t)max(0, min(255, s[k])) : 0;
    out_mask[(size_t)y * cw + x] = m ? 255 : 0;
}

void free_ws(BlendWs* ws)
{
    if (!ws) return;
    for (int i = 0; i <= kMaxBands; i++) { cudaFree(ws->dlap[i]); cudaFree(ws->dw[i]); cudaFree(ws->pyr[i]); cudaFree(ws->wp[i]); }
    delete ws;
}

inline int pad_to(int v, int nb) { return v + ((1 << nb) - v % (1 << nb)) % (1 << nb); }

struct Roi { int tlx, tly, width, height, top, left; };
Roi feed_roi(int tl_x, int tl_y, int cw, int ch, int W, int H, int nb)
{
    // MultiBandBlender::feed: gap = 3 * 2^bands, corners aligned to 2^bands, shifted back inside the canvas
    const int gap = 3 * (1 << nb);
    int tlx = tl_x - gap > 0 ? tl_x - gap : 0, tly = tl_y - gap > 0 ? tl_y - gap : 0;
    int brx = tl_x + cw + gap < W ? tl_x + cw + gap : W, bry = tl_y + ch + gap < H ? tl_y + ch + gap : H;
    tlx = (tlx >> nb) << nb; tly = (tly >> nb) << nb;
    int width = pad_to(brx - tlx, nb), height = pad_to(bry - tly, nb);
    brx = tlx + width; bry = tly + height;
    const int dy = bry - H > 0 ? bry - H : 0, dx = brx - W > 0 ? brx - W : 0;
    tlx -= dx; tly -= dy;
    Roi r; r.tlx = tlx; r.tly = tly; r.width = width; r.height = height; r.top = tl_y - tly; r.left = tl_x - tlx;
    return r;
}

}  // namespace

void uavm_blend_free(uavm_canvas* cv)
{
    if (cv && cv->blend_ws) { free_ws((BlendWs*)cv->blend_ws); cv->blend_ws = nullptr; }
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
    const int W = pad_to(cw, nb), H = pad_to(ch, nb);
    // canvas rows computed by this context: the whole padded canvas, or a band + halo (uavm_canvas_set_band).
    // Tile equivalence: cutting the pyramids at a row that is a multiple of 2^nb perturbs at most
    // 2^(nb+2) - 2 rows next to the cut after the collapse (pyrDown reaches 2 rows, pyrUp 1 row per level),
    // so a halo >= 128 rows (5 bands) leaves the band interior bit-identical to the untiled blend.
    int Y0 = 0, Y1 = H;
    if (cv->banded) {
        Y0 = cv->band_Y0; Y1 = cv->band_Y1 < H ? cv->band_Y1 : H;
        if ((Y0 % (1 << nb)) || ((Y1 % (1 << nb)) && Y1 != H)) { UAVM_SET_ERR(ctx, "band rows not aligned to 2^bands"); return UAVM_EINVAL; }
    }
    const int BH = Y1 - Y0;
    BlendWs* ws = (BlendWs*)cv->blend_ws;
    if (ws && (ws->nb != nb || ws->Y0 != Y0 || ws->Y1 != Y1)) { free_ws(ws); ws = nullptr; cv->blend_ws = nullptr; }
    auto sub_roi = [&](const ChipDesc& d, Roi& r, int& sub_t, int& sub_h) {
        r = feed_roi(d.beg_x, d.beg_y, d.chip_w, d.chip_h, W, H, nb);
        sub_t = r.tly > Y0 ? r.tly : Y0;
        const int sub_b = r.tly + r.height < Y1 ? r.tly + r.height : Y1;
        sub_h = sub_b - sub_t;
        return sub_h > 0;
    };
    if (!ws) {
        ws = new BlendWs();
        cv->blend_ws = ws;
        ws->nb = nb; ws->W = W; ws->H = H; ws->Y0 = Y0; ws->Y1 = Y1;
        size_t roi_cap = 0;
        for (int k = 0; k < cv->n; k++) {
            const ChipDesc& d = cv->desc[k];
            if (!d.keep) continue;
            Roi r; int st, sh;
            if (!sub_roi(d, r, st, sh)) continue;
            if ((size_t)r.width * sh > roi_cap) roi_cap = (size_t)r.width * sh;
        }
        ws->roi_cap = roi_cap;
        ws->lw[0] = W; ws->lh[0] = BH;
        size_t cap = roi_cap;
        for (int i = 0; i <= nb; i++) {
            if (i > 0) { ws->lw[i] = (ws->lw[i - 1] + 1) / 2; ws->lh[i] = (ws->lh[i - 1] + 1) / 2; cap = (cap + 3) / 4 + 4096; }
            const size_t px = (size_t)ws->lw[i] * ws->lh[i];
            UAVM_CUDA(ctx, cudaMalloc(&ws->dlap[i], px * 4 * sizeof(short)));
            UAVM_CUDA(ctx, cudaMalloc(&ws->dw[i], px * sizeof(float)));
            UAVM_CUDA(ctx, cudaMalloc(&ws->pyr[i], (cap + 16) * 4 * sizeof(short)));
            UAVM_CUDA(ctx, cudaMalloc(&ws->wp[i], (cap + 16) * sizeof(float)));
        }
    }
    if (cv->result_w != cw || cv->result_h != ch) {
        cudaFree(cv->d_result); cudaFree(cv->d_result_mask); cv->d_result = nullptr; cv->d_result_mask = nullptr;
        UAVM_CUDA(ctx, cudaMalloc(&cv->d_result, (size_t)cw * ch * 3));
        UAVM_CUDA(ctx, cudaMalloc(&cv->d_result_mask, (size_t)cw * ch));
        UAVM_CUDA(ctx, cudaMemsetAsync(cv->d_result, 0, (size_t)cw * ch * 3, ctx->stream));
        UAVM_CUDA(ctx, cudaMemsetAsync(cv->d_result_mask, 0, (size_t)cw * ch, ctx->stream));
        cv->result_w = cw; cv->result_h = ch;
    }
    for (int i = 0; i <= nb; i++) {
        const size_t px = (size_t)ws->lw[i] * ws->lh[i];
        UAVM_CUDA(ctx, cudaMemsetAsync(ws->dlap[i], 0, px * 4 * sizeof(short), ctx->stream));
        UAVM_CUDA(ctx, cudaMemsetAsync(ws->dw[i], 0, px * sizeof(float), ctx->stream));
    }
    // feed(), image by image in index order
    for (int k = 0; k < cv->n; k++) {
        const ChipDesc& d = cv->desc[k];
        if (!d.keep) continue;
        Roi r; int sub_t, sub_h;
        if (!sub_roi(d, r, sub_t, sub_h)) continue;
        int pw[kMaxBands + 1], ph[kMaxBands + 1];
        pw[0] = r.width; ph[0] = sub_h;
        {
            dim3 grid(((r.width + 3) / 4 + 255) / 256, sub_h);
            k7_feed_level0<<<grid, 256, 0, ctx->stream>>>(d.chip, d.chip_step, d.mask, d.mask_step, d.chip_w, d.chip_h, r.left, d.beg_y - sub_t,
                                                           r.width, sub_h, ws->pyr[0], ws->wp[0]);
            UAVM_CHECK_LAUNCH(ctx);
        }
        for (int i = 0; i < nb; i++) {
            pw[i + 1] = (pw[i] + 1) / 2; ph[i + 1] = (ph[i] + 1) / 2;
            if (pw[i + 1] >= 2 * kPdW && ph[i + 1] >= kPdH && pw[i] >= 8 && ph[i] >= 8) {
                dim3 grid((pw[i + 1] + kPdW - 1) / kPdW, (ph[i + 1] + kPdH - 1) / kPdH);
                k7_pyrdown_tiled<<<grid, dim3(kPdW, 4), 0, ctx->stream>>>(ws->pyr[i], ws->wp[i], pw[i], ph[i], ws->pyr[i + 1], ws->wp[i + 1], pw[i + 1], ph[i + 1]);
            } else {
                dim3 grid((pw[i + 1] + 255) / 256, ph[i + 1]);
                k7_pyrdown<<<grid, 256, 0, ctx->stream>>>(ws->pyr[i], ws->wp[i], pw[i], ph[i], ws->pyr[i + 1], ws->wp[i + 1], pw[i + 1], ph[i + 1]);
            }
            UAVM_CHECK_LAUNCH(ctx);
        }
        int x_tl = r.tlx, y_tl = sub_t - Y0;
        for (int i = 0; i <= nb; i++) {
            const bool quad = i < nb && !(pw[i] & 1) && !(ph[i] & 1) && !(x_tl & 1) && !(y_tl & 1) && !(ws->lw[i] & 1);
            if (quad) {
                dim3 grid((pw[i] / 2 + 255) / 256, ph[i] / 2);
                k7_lap_accumulate_quad<<<grid, 256, 0, ctx->stream>>>(ws->pyr[i], ws->pyr[i + 1], ws->wp[i], pw[i], ph[i], pw[i + 1], ph[i + 1],
                                                                       ws->dlap[i], ws->dw[i], ws->lw[i], x_tl, y_tl);
            } else {
                dim3 grid((pw[i] + 255) / 256, ph[i]);
                k7_lap_accumulate<<<grid, 256, 0, ctx->stream>>>(ws->pyr[i], i < nb ? ws->pyr[i + 1] : nullptr, ws->wp[i], pw[i], ph[i],
                                                                  i < nb ? pw[i + 1] : 0, i < nb ? ph[i + 1] : 0, ws->dlap[i], ws->dw[i], ws->lw[i], x_tl, y_tl);
            }
            UAVM_CHECK_LAUNCH(ctx);
            x_tl /= 2; y_tl /= 2;
        }
    }
    // blend(): normalise the top level, collapse from the top (normalising each level as it is consumed); the level-0 step
    // writes the cropped u8 mosaic and its mask directly
    {
        const size_t px = (size_t)ws->lw[nb] * ws->lh[nb];
        k7_normalize<<<(unsigned)((px + 255) / 256), 256, 0, ctx->stream>>>(ws->dlap[nb], ws->dw[nb], px);
        UAVM_CHECK_LAUNCH(ctx);
    }
    for (int i = nb; i > 1; i--) {
        if (!(ws->lw[i - 1] & 1) && !(ws->lh[i - 1] & 1)) {
            dim3 grid((ws->lw[i - 1] / 2 + 255) / 256, ws->lh[i - 1] / 2);
            k7_collapse_quad<<<grid, 256, 0, ctx->stream>>>(ws->dlap[i], ws->lw[i], ws->lh[i], ws->dlap[i - 1], ws->dw[i - 1], ws->lw[i - 1], ws->lh[i - 1]);
        } else {
            dim3 grid((ws->lw[i - 1] + 255) / 256, ws->lh[i - 1]);
            k7_collapse<<<grid, 256, 0, ctx->stream>>>(ws->dlap[i], ws->lw[i], ws->lh[i], ws->dlap[i - 1], ws->dw[i - 1], ws->lw[i - 1], ws->lh[i - 1]);
        }
        UAVM_CHECK_LAUNCH(ctx);
    }
    {
        const int oy0 = cv->banded ? cv->band_y0 : 0, oy1 = cv->banded ? cv->band_y1 : ch;
        dim3 grid((cw + 255) / 256, oy1 - oy0), grid4(((cw + 3) / 4 + 255) / 256, (oy1 - oy0 + 1) / 2);
        if (nb >= 1)
            k7_collapse_output<<<grid4, 256, 0, ctx->stream>>>(ws->dlap[1], ws->lw[1], ws->lh[1], ws->dlap[0], ws->dw[0], W, BH, oy0 - Y0, oy1 - Y0, cw,
                                                               cv->d_result + (size_t)oy0 * cw * 3, cv->d_result_mask + (size_t)oy0 * cw);
        else
            k7_output<<<grid, 256, 0, ctx->stream>>>(ws->dlap[0] + (size_t)(oy0 - Y0) * W * 4, ws->dw[0] + (size_t)(oy0 - Y0) * W, W, cw, oy1 - oy0,
                                                      cv->d_result + (size_t)oy0 * cw * 3, cv->d_result_mask + (size_t)oy0 * cw);
        UAVM_CHECK_LAUNCH(ctx);
    }
    cv->blended = true;
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
