// canvas.cu — device side of the warp stage of LaplacianPyramidBlending (M/MosaicImage.cpp:2205-2510):
//   K5  k5_warp_chips   bilinear inverse warp of every kept frame into its chip + validity mask (:2350-2448)
// (K6 seam masks live in masks.cu, K7 multi-band blend in blend.cu.)
//
// HBM layout: source frames stay in the caller's BGR layout (3 bytes per pixel, all frames stacked in one pool) whenever the
// frame width is a multiple of 16 — the packed affine kernel stages the BGR footprint of a chip tile with TMA and reads each
// pair of horizontally adjacent taps as three aligned shared-memory words + two funnel shifts, so no conversion pass and no
// 4th byte ever touch HBM.  Other widths use a BGRA (uchar4) pool filled by a conversion kernel (one 32-bit load per tap).
// Chips are BGRA words (row step align4(chip_w) words): B, G, R = the warped pixel, alpha != 0 = the
// reference's validity mask (:2443-2456) — the same 3 + 1 bytes per pixel as a BGR chip plus a mask plane,
// written with one coalesced 32-bit store per pixel.  The u8 mask plane (row step align4(chip_w)) is K6's output.
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "canvas.h"
#include "blend_plan.h"
#include "ptx.cuh"

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

// same, 4 pixels per thread: three aligned 32-bit loads, one 16-byte store (rows and base 4-byte aligned)
__global__ void __launch_bounds__(128) k5_bgr_to_bgra_x4(const uint8_t* __restrict__ src, int w, int step, uchar4* __restrict__ dst, int dst_step_px)
{
    const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int y = blockIdx.y;
    if (x >= w) return;
    const uint8_t* s = src + (size_t)y * step + 3 * x;
    uchar4* d = dst + (size_t)y * dst_step_px + x;
    if (x + 4 <= w && (dst_step_px % 4) == 0) {
        const uint32_t* s32 = reinterpret_cast<const uint32_t*>(s);
        const uint32_t a = __ldg(s32), b = __ldg(s32 + 1), c = __ldg(s32 + 2);        // B0G0R0B1 G1R1B2G2 R2B3G3R3
        uint4 o;
        o.x = a & 0x00ffffffu;
        o.y = __byte_perm(a, b, 0x0543) & 0x00ffffffu;
        o.z = __byte_perm(b, c, 0x0432) & 0x00ffffffu;
        o.w = c >> 8;
        *reinterpret_cast<uint4*>(d) = o;
    } else {
        for (int i = 0; i < 4 && x + i < w; i++) d[i] = make_uchar4(s[3 * i], s[3 * i + 1], s[3 * i + 2], 0);
    }
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

// Thread -> pixel mapping of both warp kernels: a CTA (32 x 8 threads) covers a 128 x 32 chip tile; a thread owns the
// pixels x = tile_x + lane + 32 k (k = 0..3) of the rows y = tile_y + ty + 8 r (r = 0..3).  The 32 lanes of a warp
// therefore sample 32 NEIGHBOURING source pixels per tap load (a 128 B line or two, + a line per source-row crossing)
// and store 32 consecutive BGRA words: the L1 sees ~3 tags per load instead of ~11 with 4 consecutive pixels per
// thread (ncu, profiles/r1d_*), which was what bounded the kernel.
constexpr int kWarpTileW = 128;   // 32 lanes x 4 px
constexpr int kWarpRows = 4;      // rows per thread
#ifndef UAVM_K5_WARPS
#define UAVM_K5_WARPS 8
#endif
constexpr int kWarpsY = UAVM_K5_WARPS;   // warps (= thread rows) per CTA; a thread's rows are kWarpsY apart
constexpr int kWarpTileH = kWarpsY * kWarpRows;
constexpr int kFpBoxW = 160, kFpBoxH = 16;        // TMA box (pixels): the staged footprint is up to kFpBoxes boxes stacked vertically
constexpr int kFpBoxBytes = kFpBoxW * kFpBoxH * 4;
constexpr int kFpBoxes = 4;                       // 160 x 64 px = 40 KB of dynamic shared memory per CTA (5 CTAs / SM)
// BGR pool: a box is 176 px x 16 rows of packed BGR (528 bytes per row = 132 u32 tensor elements; the box origin must be 16-byte
// aligned, i.e. a multiple of 16 pixels), 4 boxes = 33 KB
constexpr int kFpBoxWBgr = 176, kFpPitchBgr = kFpBoxWBgr * 3, kFpBoxBytesBgr = kFpPitchBgr * kFpBoxH;

// Generic (projective) warp, scalar.  AFFINE: inv[6] == inv[7] == 0 and inv[8] == 1, so the reference's denominator is
// exactly 1.0f for every pixel and x / 1.0f == x: the two divides per coordinate (:2359-2362) are skipped without
// changing a bit (kept as the A/B twin of the packed affine kernel below).
// BGR: the source pool holds packed BGR rows of `src_step_px` BYTES; a tap is three byte loads (this kernel is the slow path).
__device__ __forceinline__ uint32_t load_tap(const uint8_t* __restrict__ src8, size_t byte_off)
{
    return (uint32_t)__ldg(src8 + byte_off) | ((uint32_t)__ldg(src8 + byte_off + 1) << 8) | ((uint32_t)__ldg(src8 + byte_off + 2) << 16);
}

template <bool AFFINE, bool BGR>
__global__ void __launch_bounds__(32 * kWarpsY)
k5_warp_chips(const ChipDesc* __restrict__ descs, int img_w, int img_h, int src_step_px, float dgx, float dgy,
              uint32_t two23 /* = 0x4B000000, bits of 2^23: a kernel argument so PRMT takes the SELECTOR as its immediate */,
              float w1f, float h1f)
{
    const ChipDesc& D = descs[blockIdx.z];
    if (!D.keep || (D.affine != 0) != AFFINE) return;
    // chip pixels this context needs (a sharded canvas skips the tiles nothing can read)
    if ((int)(blockIdx.y * kWarpTileH) >= D.need_y1 || (int)(blockIdx.y * kWarpTileH) + kWarpTileH <= D.need_y0 ||
        (int)(blockIdx.x * kWarpTileW) >= D.need_x1 || (int)(blockIdx.x * kWarpTileW) + kWarpTileW <= D.need_x0) return;
    const int xl = blockIdx.x * kWarpTileW + threadIdx.x;
    const int ybase = blockIdx.y * kWarpTileH + threadIdx.y;
    if (blockIdx.x * kWarpTileW >= D.chip_w || blockIdx.y * kWarpTileH >= D.chip_h) return;
    const float iv0 = D.inv[0], iv1 = D.inv[1], iv2 = D.inv[2], iv3 = D.inv[3], iv4 = D.inv[4], iv5 = D.inv[5];
    const float iv6 = D.inv[6], iv7 = D.inv[7], iv8 = D.inv[8];
    const float w1 = w1f, h1 = h1f;                 // (float)(width-1), (float)(height-1): kernel arguments, not re-converted per pixel
    const uint32_t* __restrict__ src = reinterpret_cast<const uint32_t*>(D.src);
    const float fbx = (float)D.beg_x, fby = (float)D.beg_y, sx = D.sx, sy = D.sy;
    // xTemp = xDst - dGx - sx + begBoxX (:2356)
    float xt[4];
#pragma unroll
    for (int i = 0; i < 4; i++) xt[i] = (float)(xl + 32 * i) - dgx - sx + fbx;
#pragma unroll
    for (int ry = 0; ry < kWarpRows; ry++) {
        const int yd = ybase + ry * kWarpsY;
        if (yd >= D.chip_h) break;
        const float yt = (float)yd - dgy - sy + fby;               // yTemp (:2357)
        const float ya = yt * iv1, yb = yt * iv4;
        uint32_t* crow = D.chip + (size_t)yd * D.chip_step;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            if (xl + 32 * i >= D.chip_w) break;
            float xs = xt[i] * iv0 + ya + iv2;
            float ys = xt[i] * iv3 + yb + iv5;
            if (!AFFINE) {
                const float den = xt[i] * iv6 + yt * iv7 + iv8;
                xs = xs / den;
                ys = ys / den;
            }
            uint32_t word = 0;                                     // outside the source: BGR = 0, alpha = 0 (mask 0, :2443-2446)
            if ((xs >= 0.0f) && (xs < w1) && (ys >= 0.0f) && (ys < h1)) {
                const int iy = __float2int_rz(ys), ix = __float2int_rz(xs);
                const float p = ys - (float)iy, q = xs - (float)ix;
                const float omp = 1.0f - p, omq = 1.0f - q;
                const float c_omp = -8388608.0f * omp, c_p = -8388608.0f * p;        // exact (power-of-two scale)
                uint32_t t00, t01, t10, t11;
                if (BGR) {
                    const uint8_t* s8 = reinterpret_cast<const uint8_t*>(D.src);
                    const size_t o = (size_t)iy * (size_t)src_step_px + 3u * (size_t)ix;
                    t00 = load_tap(s8, o); t01 = load_tap(s8, o + 3); t10 = load_tap(s8, o + src_step_px); t11 = load_tap(s8, o + src_step_px + 3);
                } else {
                    const uint32_t off = (uint32_t)(iy * src_step_px + ix);
                    t00 = __ldg(src + off); t01 = __ldg(src + off + 1u);
                    t10 = __ldg(src + off + (uint32_t)src_step_px); t11 = __ldg(src + off + (uint32_t)src_step_px + 1u);
                }
                const uint32_t b = bilinear_channel<0>(t00, t01, t10, t11, p, q, omp, omq, c_omp, c_p, two23);
                const uint32_t g = bilinear_channel<1>(t00, t01, t10, t11, p, q, omp, omq, c_omp, c_p, two23);
                const uint32_t r = bilinear_channel<2>(t00, t01, t10, t11, p, q, omp, omq, c_omp, c_p, two23);
                word = __byte_perm(__byte_perm(b, g, 0x0040), r, 0x0410) | 0xff000000u;
            }
            crow[xl + 32 * i] = word;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Packed-f32x2 variant of the AFFINE warp.  sm_100a has two-lane FP32 instructions (FFMA2 / FMUL2 / FADD2):
// one issue slot performs the same IEEE-754 round-to-nearest operation on two independent floats, so the
// kernel evaluates the reference expression for two pixels per instruction.  Every lane still performs exactly
// the reference's operation sequence:
//   * a product that feeds a sum must not be contracted.  ptxas contracts mul.f32x2 + add.f32x2 into FFMA2
//     even under -fmad=false, so those sums are written as fma2(t, ONE, s) with ONE = 1.0f passed as a KERNEL
//     ARGUMENT (opaque to ptxas): fma(t, 1, s) rounds t + s once, i.e. it IS the IEEE addition;
//   * int(v) for 0 <= v < 2^23 is add.rz(v, 2^23): the sum is rounded toward zero to a multiple of 1.0, which
//     leaves 2^23 + floor(v) — the integer sits in the low mantissa bits, and subtracting 2^23 (exact) gives
//     (float)int(v).  This removes F2I / I2F (XU pipe, 1/8 rate) from the kernel altogether.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpk2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ void unpk2u(f32x2 v, uint32_t& lo, uint32_t& hi) { asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) { f32x2 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 add2_rz(f32x2 a, f32x2 b) { f32x2 d; asm("add.rz.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

// one channel, two pixels: returns 2^23 + int(v) per lane (the output byte is the low byte of each word)
template <int K>
__device__ __forceinline__ f32x2 bilinear_channel2(const uint32_t (&ta)[4], const uint32_t (&tb)[4], f32x2 P, f32x2 Q, f32x2 OMP, f32x2 OMQ,
                                                   f32x2 C_OMP, f32x2 C_P, f32x2 ONE, f32x2 T23, uint32_t two23)
{
    const f32x2 G1 = pk2(__int_as_float(__byte_perm(ta[0], two23, 0x7540 | K)), __int_as_float(__byte_perm(tb[0], two23, 0x7540 | K)));
    const f32x2 G2 = pk2(__int_as_float(__byte_perm(ta[1], two23, 0x7540 | K)), __int_as_float(__byte_perm(tb[1], two23, 0x7540 | K)));
    const f32x2 G3 = pk2(__int_as_float(__byte_perm(ta[2], two23, 0x7540 | K)), __int_as_float(__byte_perm(tb[2], two23, 0x7540 | K)));
    const f32x2 G4 = pk2(__int_as_float(__byte_perm(ta[3], two23, 0x7540 | K)), __int_as_float(__byte_perm(tb[3], two23, 0x7540 | K)));
    const f32x2 a1 = fma2(G1, OMP, C_OMP), a2 = fma2(G2, OMP, C_OMP);          // g1*(1-p), g2*(1-p)   (exact-FFMA trick, see bilinear_channel)
    const f32x2 a3 = fma2(G3, P, C_P), a4 = fma2(G4, P, C_P);                  // g3*p, g4*p
    f32x2 s = mul2(a1, OMQ);
    s = fma2(mul2(a2, Q), ONE, s);                                             // + (g2*(1-p))*q      (ONE is opaque: no contraction)
    s = fma2(mul2(a3, OMQ), ONE, s);
    s = fma2(mul2(a4, Q), ONE, s);
    return add2_rz(s, T23);                                                    // 2^23 + int(v)
}

// BGR taps: f0 | f1 hold the 8 bytes that start at the first tap of a row (B0 G0 R0 B1 | G1 R1 ..), so the left tap's channel K
// is byte K of f0 and the right tap's is byte K + 3 of the pair: still ONE PRMT per tap and channel.
template <int K>
__device__ __forceinline__ f32x2 bilinear_channel2_bgr(uint32_t f0a, uint32_t f1a, uint32_t g0a, uint32_t g1a, uint32_t f0b, uint32_t f1b, uint32_t g0b, uint32_t g1b,
                                                       f32x2 P, f32x2 Q, f32x2 OMP, f32x2 OMQ, f32x2 C_OMP, f32x2 C_P, f32x2 ONE, f32x2 T23, uint32_t two23)
{
    constexpr uint32_t kSelR = 0x7540u | (K == 0 ? 3u : (uint32_t)(K - 1));      // right tap: byte K + 3 of (f0, f1)
    const f32x2 G1 = pk2(__int_as_float(__byte_perm(f0a, two23, 0x7540 | K)), __int_as_float(__byte_perm(f0b, two23, 0x7540 | K)));
    const f32x2 G2 = pk2(__int_as_float(__byte_perm(K == 0 ? f0a : f1a, two23, kSelR)), __int_as_float(__byte_perm(K == 0 ? f0b : f1b, two23, kSelR)));
    const f32x2 G3 = pk2(__int_as_float(__byte_perm(g0a, two23, 0x7540 | K)), __int_as_float(__byte_perm(g0b, two23, 0x7540 | K)));
    const f32x2 G4 = pk2(__int_as_float(__byte_perm(K == 0 ? g0a : g1a, two23, kSelR)), __int_as_float(__byte_perm(K == 0 ? g0b : g1b, two23, kSelR)));
    const f32x2 a1 = fma2(G1, OMP, C_OMP), a2 = fma2(G2, OMP, C_OMP);
    const f32x2 a3 = fma2(G3, P, C_P), a4 = fma2(G4, P, C_P);
    f32x2 s = mul2(a1, OMQ);
    s = fma2(mul2(a2, Q), ONE, s);
    s = fma2(mul2(a3, OMQ), ONE, s);
    s = fma2(mul2(a4, Q), ONE, s);
    return add2_rz(s, T23);
}

// Source staging (SMEM = true).  The taps of a 128 x 32 chip tile fall into a small source rectangle (its footprint:
// the image of the tile under the inverse transform, ~140 x 47 px for a UAV strip).  Fetching taps with per-thread
// loads leaves every warp waiting on an L2/HBM round trip per pixel pair (ncu: ~90 % of resident warps stalled on the
// long scoreboard, no unit above 60 %).  Instead warp 0 copies the footprint rows into shared memory with bulk async
// copies (cp.async.bulk -> UBLKCP, completion on an mbarrier) while all threads do their per-thread set-up, and the
// taps become LDS: with lane-consecutive pixels and a row pitch that is a multiple of 32 words they are bank-conflict
// free.  Tiles whose footprint exceeds the shared-memory budget use the direct global-load path (SMEM = false).
//
// CHECK = false: the caller proved every pixel of this thread's 4 x 4 block lies inside the source, so the
// per-pixel validity test, the tap clamping and the zeroing are skipped.

__device__ __forceinline__ uint32_t lds_u32(const uint8_t* base, uint32_t off) { return *reinterpret_cast<const uint32_t*>(base + off); }

template <bool CHECK, bool SMEM, bool BGR>
__device__ __forceinline__ void warp_affine_rows(const uint8_t* fp_base, const ChipDesc& D, const f32x2 (&MX)[2], const f32x2 (&MY)[2], int xl, int ybase, float dgy, float sy, float fby,
                                                 float iv1, float iv4, f32x2 IV2, f32x2 IV5, uint32_t step4, unsigned long long base_adj,
                                                 uint32_t sm_adj, float clampx, float clampy,
                                                 uint32_t one_u, uint32_t two23, float w1f, float h1f, f32x2 ONE)
{
    const f32x2 T23 = pk2(8388608.0f, 8388608.0f), N23 = pk2(-8388608.0f, -8388608.0f), ONEI = pk2(1.0f, 1.0f);
    constexpr uint32_t sm_pitch4 = BGR ? (uint32_t)kFpPitchBgr : kFpBoxW * 4u;
#pragma unroll
    for (int ry = 0; ry < kWarpRows; ry++) {
        const int yd = ybase + ry * kWarpsY;
        if (CHECK && yd >= D.chip_h) break;
        const float yt = ((float)ybase + (float)(ry * kWarpsY)) - dgy - sy + fby;       // yTemp (:2357); (float)(ybase + kWarpsY ry) is exact
        const float ya = yt * iv1, yb = yt * iv4;
        const f32x2 YA = pk2(ya, ya), YB = pk2(yb, yb);
        uint32_t* crow = D.chip + (size_t)yd * D.chip_step + xl;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            f32x2 XS = add2(fma2(MX[h], ONE, YA), IV2);            // (x*inv0 + y*inv1) + inv2
            f32x2 YS = add2(fma2(MY[h], ONE, YB), IV5);
            bool va = true, vb = true;
            if (CHECK) {
                float xsa, xsb, ysa, ysb;
                unpk2(XS, xsa, xsb); unpk2(YS, ysa, ysb);
                va = (xsa >= 0.0f) && (xsa < w1f) && (ysa >= 0.0f) && (ysa < h1f);
                vb = (xsb >= 0.0f) && (xsb < w1f) && (ysb >= 0.0f) && (ysb < h1f);
                if (!va) { xsa = clampx; ysa = clampy; }            // keep the taps in bounds; the word is zeroed below
                if (!vb) { xsb = clampx; ysb = clampy; }
                XS = pk2(xsa, xsb); YS = pk2(ysa, ysb);
            }
            const f32x2 TX = add2_rz(XS, T23), TY = add2_rz(YS, T23);      // 2^23 + int(xs), 2^23 + int(ys)
            const f32x2 Q = sub2(XS, sub2(TX, T23)), P = sub2(YS, sub2(TY, T23));   // q = xs - ix, p = ys - iy
            const f32x2 OMP = sub2(ONEI, P), OMQ = sub2(ONEI, Q);
            const f32x2 C_OMP = mul2(OMP, N23), C_P = mul2(P, N23);        // -2^23 * w, exact
            uint32_t ixa, ixb, iya, iyb;
            unpk2u(TX, ixa, ixb); unpk2u(TY, iya, iyb);
            if (BGR) {
                // byte offset of tap (ix, iy) = adj + iy_bits * pitch + ix_bits * 3; the 6 bytes of a tap pair come from three
                // aligned words and two funnel shifts (pitches are multiples of 4, so both rows share the shift)
                uint32_t f0a, f1a, g0a, g1a, f0b, f1b, g0b, g1b;
                if (SMEM) {
                    const uint32_t oa = iya * sm_pitch4 + (ixa * 3u + sm_adj), ob = iyb * sm_pitch4 + (ixb * 3u + sm_adj);
                    const uint32_t ba = oa & ~3u, bb = ob & ~3u, sa = (oa & 3u) * 8u, sb = (ob & 3u) * 8u;
                    const uint32_t a0 = lds_u32(fp_base, ba), a1 = lds_u32(fp_base, ba + 4u), a2 = lds_u32(fp_base, ba + 8u);
                    const uint32_t a3 = lds_u32(fp_base, ba + sm_pitch4), a4 = lds_u32(fp_base, ba + sm_pitch4 + 4u), a5 = lds_u32(fp_base, ba + sm_pitch4 + 8u);
                    const uint32_t b0 = lds_u32(fp_base, bb), b1 = lds_u32(fp_base, bb + 4u), b2 = lds_u32(fp_base, bb + 8u);
                    const uint32_t b3 = lds_u32(fp_base, bb + sm_pitch4), b4 = lds_u32(fp_base, bb + sm_pitch4 + 4u), b5 = lds_u32(fp_base, bb + sm_pitch4 + 8u);
                    f0a = __funnelshift_r(a0, a1, sa); f1a = __funnelshift_r(a1, a2, sa); g0a = __funnelshift_r(a3, a4, sa); g1a = __funnelshift_r(a4, a5, sa);
                    f0b = __funnelshift_r(b0, b1, sb); f1b = __funnelshift_r(b1, b2, sb); g0b = __funnelshift_r(b3, b4, sb); g1b = __funnelshift_r(b4, b5, sb);
                } else {
                    const unsigned long long pa = base_adj + (unsigned long long)iya * step4 + (unsigned long long)ixa * 3u;
                    const unsigned long long pb = base_adj + (unsigned long long)iyb * step4 + (unsigned long long)ixb * 3u;
                    const unsigned long long qa = pa & ~3ull, qb = pb & ~3ull;
                    const uint32_t sa = ((uint32_t)pa & 3u) * 8u, sb = ((uint32_t)pb & 3u) * 8u;
                    const unsigned long long qa1 = qa + (unsigned long long)step4 * one_u, qb1 = qb + (unsigned long long)step4 * one_u;
                    const uint32_t a0 = __ldg(reinterpret_cast<const uint32_t*>(qa)), a1 = __ldg(reinterpret_cast<const uint32_t*>(qa + 4)), a2 = __ldg(reinterpret_cast<const uint32_t*>(qa + 8));
                    const uint32_t a3 = __ldg(reinterpret_cast<const uint32_t*>(qa1)), a4 = __ldg(reinterpret_cast<const uint32_t*>(qa1 + 4)), a5 = __ldg(reinterpret_cast<const uint32_t*>(qa1 + 8));
                    const uint32_t b0 = __ldg(reinterpret_cast<const uint32_t*>(qb)), b1 = __ldg(reinterpret_cast<const uint32_t*>(qb + 4)), b2 = __ldg(reinterpret_cast<const uint32_t*>(qb + 8));
                    const uint32_t b3 = __ldg(reinterpret_cast<const uint32_t*>(qb1)), b4 = __ldg(reinterpret_cast<const uint32_t*>(qb1 + 4)), b5 = __ldg(reinterpret_cast<const uint32_t*>(qb1 + 8));
                    f0a = __funnelshift_r(a0, a1, sa); f1a = __funnelshift_r(a1, a2, sa); g0a = __funnelshift_r(a3, a4, sa); g1a = __funnelshift_r(a4, a5, sa);
                    f0b = __funnelshift_r(b0, b1, sb); f1b = __funnelshift_r(b1, b2, sb); g0b = __funnelshift_r(b3, b4, sb); g1b = __funnelshift_r(b4, b5, sb);
                }
                uint32_t ba_, bb_, ga_, gb_, ra_, rb_;
                unpk2u(bilinear_channel2_bgr<0>(f0a, f1a, g0a, g1a, f0b, f1b, g0b, g1b, P, Q, OMP, OMQ, C_OMP, C_P, ONE, T23, two23), ba_, bb_);
                unpk2u(bilinear_channel2_bgr<1>(f0a, f1a, g0a, g1a, f0b, f1b, g0b, g1b, P, Q, OMP, OMQ, C_OMP, C_P, ONE, T23, two23), ga_, gb_);
                unpk2u(bilinear_channel2_bgr<2>(f0a, f1a, g0a, g1a, f0b, f1b, g0b, g1b, P, Q, OMP, OMQ, C_OMP, C_P, ONE, T23, two23), ra_, rb_);
                uint32_t wa = __byte_perm(__byte_perm(ba_, ga_, 0x0040), ra_, 0x7410);
                uint32_t wb = __byte_perm(__byte_perm(bb_, gb_, 0x0040), rb_, 0x7410);
                if (CHECK) {
                    if (!va) wa = 0u;
                    if (!vb) wb = 0u;
                    if (xl + 64 * h < D.chip_w) crow[64 * h] = wa;
                    if (xl + 64 * h + 32 < D.chip_w) crow[64 * h + 32] = wb;
                } else {
                    crow[64 * h] = wa; crow[64 * h + 32] = wb;
                }
                continue;
            }
            uint32_t ta[4], tb[4];
            if (SMEM) {
                // offset of tap (ix, iy) in the staged footprint = sm_adj + iy_bits * pitch + ix_bits * 4 (biases and footprint origin folded into sm_adj)
                const uint32_t sa = iya * sm_pitch4 + (ixa * 4u + sm_adj), sb = iyb * sm_pitch4 + (ixb * 4u + sm_adj);
                ta[0] = lds_u32(fp_base, sa); ta[1] = lds_u32(fp_base, sa + 4u); ta[2] = lds_u32(fp_base, sa + sm_pitch4); ta[3] = lds_u32(fp_base, sa + sm_pitch4 + 4u);
                tb[0] = lds_u32(fp_base, sb); tb[1] = lds_u32(fp_base, sb + 4u); tb[2] = lds_u32(fp_base, sb + sm_pitch4); tb[3] = lds_u32(fp_base, sb + sm_pitch4 + 4u);
            } else {
                // byte address of tap (ix, iy) = base_adj + iy_bits * 4 step + ix_bits * 4 (the 0x4B000000 biases are folded into base_adj)
                const unsigned long long pa = base_adj + (unsigned long long)iya * step4 + (unsigned long long)ixa * 4u;
                const unsigned long long pb = base_adj + (unsigned long long)iyb * step4 + (unsigned long long)ixb * 4u;
                const unsigned long long pa1 = pa + (unsigned long long)step4 * one_u, pb1 = pb + (unsigned long long)step4 * one_u;   // one IMAD.WIDE each
                ta[0] = __ldg(reinterpret_cast<const uint32_t*>(pa)); ta[1] = __ldg(reinterpret_cast<const uint32_t*>(pa + 4));
                ta[2] = __ldg(reinterpret_cast<const uint32_t*>(pa1)); ta[3] = __ldg(reinterpret_cast<const uint32_t*>(pa1 + 4));
                tb[0] = __ldg(reinterpret_cast<const uint32_t*>(pb)); tb[1] = __ldg(reinterpret_cast<const uint32_t*>(pb + 4));
                tb[2] = __ldg(reinterpret_cast<const uint32_t*>(pb1)); tb[3] = __ldg(reinterpret_cast<const uint32_t*>(pb1 + 4));
            }
            uint32_t ba, bb, ga, gb, ra, rb;
            unpk2u(bilinear_channel2<0>(ta, tb, P, Q, OMP, OMQ, C_OMP, C_P, ONE, T23, two23), ba, bb);
            unpk2u(bilinear_channel2<1>(ta, tb, P, Q, OMP, OMQ, C_OMP, C_P, ONE, T23, two23), ga, gb);
            unpk2u(bilinear_channel2<2>(ta, tb, P, Q, OMP, OMQ, C_OMP, C_P, ONE, T23, two23), ra, rb);
            // BGRA word: B, G, R = low bytes; alpha = top byte of 2^23 + r (0x4B, non-zero = valid)
            uint32_t wa = __byte_perm(__byte_perm(ba, ga, 0x0040), ra, 0x7410);
            uint32_t wb = __byte_perm(__byte_perm(bb, gb, 0x0040), rb, 0x7410);
            if (CHECK) {
                if (!va) wa = 0u;
                if (!vb) wb = 0u;
                if (xl + 64 * h < D.chip_w) crow[64 * h] = wa;
                if (xl + 64 * h + 32 < D.chip_w) crow[64 * h + 32] = wb;
            } else {
                crow[64 * h] = wa; crow[64 * h + 32] = wb;
            }
        }
    }
}

// Footprint of chip tile (bx, by): the source rectangle its taps fall into.  xs, ys are monotone in the pixel column and
// in the row (every rounding step is monotone), so their extremes over the tile are attained at its corner pixels.
// Returns 0: use direct loads, 1: stage rows [y0, y0 + rows) x columns [x0, x0 + kFpBoxW), 2: tile entirely outside.
template <bool BGR>
__device__ __forceinline__ int tile_footprint(const ChipDesc& D, int bx, int by, int img_w, int img_h, float dgx, float dgy,
                                              float w1f, float h1f, int max_boxes, int& x0, int& y0, int& rows)
{
    const float iv0 = D.inv[0], iv1 = D.inv[1], iv2 = D.inv[2], iv3 = D.inv[3], iv4 = D.inv[4], iv5 = D.inv[5];
    const float fbx = (float)D.beg_x, fby = (float)D.beg_y, sx = D.sx, sy = D.sy;
    const int cx0 = bx * kWarpTileW, cy0 = by * kWarpTileH;
    const int cx1 = min(cx0 + kWarpTileW, D.chip_w) - 1, cy1 = min(cy0 + kWarpTileH, D.chip_h) - 1;
    float xmn = 3.0e38f, xmx = -3.0e38f, ymn = 3.0e38f, ymx = -3.0e38f;
#pragma unroll
    for (int c = 0; c < 4; c++) {
        const float xt = (float)((c & 1) ? cx1 : cx0) - dgx - sx + fbx, yt = (float)((c & 2) ? cy1 : cy0) - dgy - sy + fby;
        const float xs = xt * iv0 + yt * iv1 + iv2, ys = xt * iv3 + yt * iv4 + iv5;
        xmn = fminf(xmn, xs); xmx = fmaxf(xmx, xs); ymn = fminf(ymn, ys); ymx = fmaxf(ymx, ys);
    }
    if (!(xmx >= 0.0f) || !(xmn < w1f) || !(ymx >= 0.0f) || !(ymn < h1f)) return 2;
    // valid samples have 0 <= xs < w-1: taps in columns int(xs), int(xs)+1 <= w-1 (same for rows)
    x0 = ((int)fmaxf(xmn, 0.0f)) & (BGR ? ~15 : ~3);                           // 16-byte aligned box origin
    const int x1 = min((int)fminf(xmx, w1f) + 1, img_w - 1);
    y0 = (int)fmaxf(ymn, 0.0f);
    const int y1 = min((int)fminf(ymx, h1f) + 1, img_h - 1);
    rows = y1 - y0 + 1;
    // BGR: a tap pair reads the 12 bytes from the word holding its first byte, i.e. up to 2 pixels past the right tap
    return (x1 - x0 + 1 + (BGR ? 2 : 0) <= (BGR ? kFpBoxWBgr : kFpBoxW) && rows <= max_boxes * kFpBoxH) ? 1 : 0;
}

template <bool BGR>
__global__ void __launch_bounds__(32 * kWarpsY, 1024 / (32 * kWarpsY))
k5_warp_affine_x2(const __grid_constant__ CUtensorMap tmap_src, const ChipDesc* __restrict__ descs, int img_w, int img_h, int src_step_px, float dgx, float dgy,
                  uint32_t two23, float w1f, float h1f, float one,
                  uint32_t one_u /* = 1, opaque: row-1 tap address = row-0 address + step4 * 1 as a single 64-bit IMAD */,
                  unsigned long long bias /* = 0x4B000000 * (row bytes + pixel bytes): the add.rz biases of iy and ix in byte-address units */,
                  int max_boxes /* TMA boxes of dynamic shared memory for the source footprint; 0 = always load taps directly */)
{
    extern __shared__ __align__(128) uint8_t fp_smem[];
    __shared__ __align__(8) uint64_t fp_bar;
    __shared__ int fp[4];                            // footprint: x0, y0, rows, state

    const ChipDesc& D = descs[blockIdx.z];
    if (!D.keep || !D.affine) return;
    if ((int)(blockIdx.y * kWarpTileH) >= D.need_y1 || (int)(blockIdx.y * kWarpTileH) + kWarpTileH <= D.need_y0 ||
        (int)(blockIdx.x * kWarpTileW) >= D.need_x1 || (int)(blockIdx.x * kWarpTileW) + kWarpTileW <= D.need_x0) return;
    const int xl = blockIdx.x * kWarpTileW + threadIdx.x;
    const int ybase = blockIdx.y * kWarpTileH + threadIdx.y;
    if (blockIdx.x * kWarpTileW >= D.chip_w || blockIdx.y * kWarpTileH >= D.chip_h) return;
    const float iv0 = D.inv[0], iv1 = D.inv[1], iv2 = D.inv[2], iv3 = D.inv[3], iv4 = D.inv[4], iv5 = D.inv[5];
    const float fbx = (float)D.beg_x, fby = (float)D.beg_y, sx = D.sx, sy = D.sy;
    const int tid = threadIdx.y * 32 + threadIdx.x;

    // ---- thread 0: footprint of this tile, then one TMA box copy per 16 footprint rows (completion on fp_bar)
    if (tid == 0) {
        int state = 0;                                // 0: direct loads, 1: staged, 2: tile entirely outside the source
        if (max_boxes > 0) {
            int x0, y0, rows;
            state = tile_footprint<BGR>(D, blockIdx.x, blockIdx.y, img_w, img_h, dgx, dgy, w1f, h1f, max_boxes, x0, y0, rows);
            if (state == 1) {
                fp[0] = x0; fp[1] = y0; fp[2] = rows;
                uavm::ptx::mbar_init(&fp_bar, 1); uavm::ptx::fence_mbar_init();
                const int boxes = (rows + kFpBoxH - 1) / kFpBoxH;
                constexpr int box_bytes = BGR ? kFpBoxBytesBgr : kFpBoxBytes;
                uavm::ptx::mbar_arrive_expect_tx(&fp_bar, (uint32_t)(boxes * box_bytes));
                for (int k = 0; k < boxes; k++)       // BGR: the tensor's elements are u32 words of the packed rows (x0 is a multiple of 16 px = 12 words)
                    uavm::ptx::tma_load_2d(fp_smem + k * box_bytes, &tmap_src, &fp_bar, BGR ? (x0 * 3) / 4 : x0, D.src_row0 + y0 + k * kFpBoxH);
            }
        }
        fp[3] = state;
    }
    // per-thread set-up first: it overlaps thread 0's footprint plan and the flight of the TMA copies
    const f32x2 ONE = pk2(one, one);
    const f32x2 IV2 = pk2(iv2, iv2), IV5 = pk2(iv5, iv5);
    const uint32_t step4 = BGR ? (uint32_t)src_step_px : 4u * (uint32_t)src_step_px;      // source row pitch in bytes (BGR: src_step_px is already bytes)
    // the raw add.rz results are 0x4B000000 + ix / + iy: fold both biases into the frame base (64-bit, wraps are harmless)
    const unsigned long long base_adj = reinterpret_cast<unsigned long long>(D.src) - bias;
    f32x2 MX[2], MY[2];                              // xTemp * inv[0], xTemp * inv[3] for the pixel pairs (k = 2h, 2h + 1): row invariant
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const float xa = (float)(xl + 64 * h) - dgx - sx + fbx, xb = (float)(xl + 64 * h + 32) - dgx - sx + fbx;   // :2356
        MX[h] = pk2(xa * iv0, xb * iv0); MY[h] = pk2(xa * iv3, xb * iv3);
    }
    // Is the whole 4 x 4 block of this thread inside the source?  (same monotonicity argument, per thread)
    bool inside = (ybase + kWarpsY * (kWarpRows - 1) < D.chip_h) && (xl + 96 < D.chip_w);
    {
        float mxa, mxb, mya, myb, t;
        unpk2(MX[0], mxa, t); unpk2(MX[1], t, mxb); unpk2(MY[0], mya, t); unpk2(MY[1], t, myb);
#pragma unroll
        for (int c = 0; c < 2; c++) {
            const float yt = (float)(ybase + c * kWarpsY * (kWarpRows - 1)) - dgy - sy + fby;
            const float ya = yt * iv1, yb = yt * iv4;
            const float xs0 = mxa + ya + iv2, xs1 = mxb + ya + iv2, ys0 = mya + yb + iv5, ys1 = myb + yb + iv5;
            inside = inside && (xs0 >= 0.0f) && (xs0 < w1f) && (xs1 >= 0.0f) && (xs1 < w1f) && (ys0 >= 0.0f) && (ys0 < h1f) && (ys1 >= 0.0f) && (ys1 < h1f);
        }
    }
    __syncthreads();
    const int state = fp[3];
    if (state == 2) {                                 // nothing of this tile is inside the source: BGR = 0, alpha = 0
#pragma unroll
        for (int ry = 0; ry < kWarpRows; ry++) {
            const int yd = ybase + ry * kWarpsY;
            if (yd >= D.chip_h) break;
#pragma unroll
            for (int i = 0; i < 4; i++) if (xl + 32 * i < D.chip_w) D.chip[(size_t)yd * D.chip_step + xl + 32 * i] = 0u;
        }
        return;
    }
    uint32_t sm_adj = 0;
    constexpr uint32_t sm_pitch4 = BGR ? (uint32_t)kFpPitchBgr : kFpBoxW * 4u;      // the boxes stack into rows of one pitch
    constexpr uint32_t px_bytes = BGR ? 3u : 4u;
    float clampx = 0.0f, clampy = 0.0f;
    if (state == 1) {
        const int x0 = fp[0], y0 = fp[1];
        sm_adj = 0u - (0x4B000000u * sm_pitch4 + 0x4B000000u * px_bytes) - ((uint32_t)y0 * sm_pitch4 + (uint32_t)x0 * px_bytes);   // offset from fp_smem
        clampx = (float)x0; clampy = (float)y0;
    }
#define K5_ROWS_ARGS fp_smem, D, MX, MY, xl, ybase, dgy, sy, fby, iv1, iv4, IV2, IV5, step4, base_adj, sm_adj, clampx, clampy, one_u, two23, w1f, h1f, ONE
    if (state == 1) {
        uavm::ptx::mbar_wait(&fp_bar, 0);             // footprint has landed in shared memory
        if (inside) warp_affine_rows<false, true, BGR>(K5_ROWS_ARGS);
        else        warp_affine_rows<true, true, BGR>(K5_ROWS_ARGS);
    } else {
        if (inside) warp_affine_rows<false, false, BGR>(K5_ROWS_ARGS);
        else        warp_affine_rows<true, false, BGR>(K5_ROWS_ARGS);
    }
#undef K5_ROWS_ARGS
}

// chip (BGRA, alpha != 0 = inside the source) -> the u8 mask plane (255 / 0): the validity mask the reference keeps
// per chip (:2443-2456), materialised only when somebody needs it before K6 overwrites it with the seam masks
__global__ void __launch_bounds__(256) k5_alpha_to_mask(const ChipDesc* __restrict__ descs)
{
    const ChipDesc& D = descs[blockIdx.z];
    if (!D.keep) return;
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4, y = blockIdx.y;
    if (x0 >= D.chip_w || y >= D.chip_h) return;
    const uint4 v = *reinterpret_cast<const uint4*>(D.chip + (size_t)y * D.chip_step + x0);
    const uint32_t bits = ((v.x >> 24) ? 0xffu : 0u) | ((v.y >> 24) ? 0xff00u : 0u) | ((v.z >> 24) ? 0xff0000u : 0u) | ((v.w >> 24) ? 0xff000000u : 0u);
    *reinterpret_cast<uint32_t*>(D.mask + (size_t)y * D.mask_step + x0) = bits;
}

// BGRA chip -> packed BGR rows (for uavm_canvas_get_chip)
__global__ void __launch_bounds__(256) k5_chip_to_bgr(const uint32_t* __restrict__ chip, int chip_step, int w, int h, uint8_t* __restrict__ out)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w || y >= h) return;
    const uint32_t v = chip[(size_t)y * chip_step + x];
    uint8_t* o = out + ((size_t)y * w + x) * 3;
    o[0] = (uint8_t)v; o[1] = (uint8_t)(v >> 8); o[2] = (uint8_t)(v >> 16);
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
    cv->n = n_images; cv->img_w = img_w; cv->img_h = img_h;
    cv->src_bgr = (img_w % 16 == 0) && getenv("UAVM_K5_BGRA") == nullptr;      // UAVM_K5_BGRA: force the BGRA pool (A/B measurements)
    cv->src_step_px = cv->src_bgr ? 3 * img_w : img_w;                         // BGR pool: row pitch in BYTES
    cv->last_warp_ev.assign(n_images, -1);
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
        d.mask_step = align4(c.chip_w); d.chip_step = d.mask_step;
        d.beg_x = c.beg_x; d.beg_y = c.beg_y; d.sx = c.sx; d.sy = c.sy;
        memcpy(d.inv, c.inv, sizeof(d.inv)); memcpy(d.quad, c.quad, sizeof(d.quad));
        d.affine = (c.inv[6] == 0.0f && c.inv[7] == 0.0f && c.inv[8] == 1.0f) ? 1 : 0;
        d.need_x0 = 0; d.need_y0 = 0; d.need_x1 = c.chip_w; d.need_y1 = c.chip_h;
        coff[k] = chip_off; moff[k] = mask_off;
        chip_off += (size_t)d.chip_step * d.chip_h * 4; mask_off += (size_t)d.mask_step * d.chip_h;
        chip_off = (chip_off + 255) & ~(size_t)255; mask_off = (mask_off + 255) & ~(size_t)255;
        if (c.chip_w > cv->max_chip_w) cv->max_chip_w = c.chip_w;
        if (c.chip_h > cv->max_chip_h) cv->max_chip_h = c.chip_h;
    }
    cv->chips_bytes = chip_off; cv->masks_bytes = mask_off;
    UAVM_CUDA_OR(ctx, cudaSetDevice(ctx->device), uavm_canvas_destroy(ctx, cv));
    UAVM_CUDA_OR(ctx, cudaMalloc(&cv->d_src, (size_t)n_images * img_h * cv->src_step_px * (cv->src_bgr ? 1 : sizeof(uchar4)) + 256), uavm_canvas_destroy(ctx, cv));
    UAVM_CUDA_OR(ctx, cudaMalloc(&cv->d_chips, chip_off + 256), uavm_canvas_destroy(ctx, cv));
    UAVM_CUDA_OR(ctx, cudaMalloc(&cv->d_masks, mask_off + 256), uavm_canvas_destroy(ctx, cv));
    UAVM_CUDA_OR(ctx, cudaMemsetAsync(cv->d_chips, 0, chip_off + 256, ctx->stream), uavm_canvas_destroy(ctx, cv));       // row padding stays zero
    UAVM_CUDA_OR(ctx, cudaMemsetAsync(cv->d_masks, 0, mask_off + 256, ctx->stream), uavm_canvas_destroy(ctx, cv));
    UAVM_CUDA_OR(ctx, cudaMalloc(&cv->d_desc, (size_t)n_images * sizeof(ChipDesc)), uavm_canvas_destroy(ctx, cv));
    cv->stage_pitch = (img_w * 3 + 15) & ~15;
    cv->stage_bytes = (size_t)img_h * cv->stage_pitch;
    for (int s = 0; s < uavm_canvas::kStageSlots; s++) {
        if (!cv->src_bgr) UAVM_CUDA_OR(ctx, cudaMalloc(&cv->d_stage[s], cv->stage_bytes), uavm_canvas_destroy(ctx, cv));
        UAVM_CUDA_OR(ctx, cudaEventCreateWithFlags(&cv->ev_copied[s], cudaEventDisableTiming), uavm_canvas_destroy(ctx, cv));
        UAVM_CUDA_OR(ctx, cudaEventCreateWithFlags(&cv->ev_free[s], cudaEventDisableTiming), uavm_canvas_destroy(ctx, cv));
    }
    if (cv->src_bgr)
        for (int e = 0; e < uavm_canvas::kWarpEvents; e++) UAVM_CUDA_OR(ctx, cudaEventCreateWithFlags(&cv->ev_warp[e], cudaEventDisableTiming), uavm_canvas_destroy(ctx, cv));
    for (int k = 0; k < n_images; k++) {
        ChipDesc& d = cv->desc[k];
        d.src = cv->src_bgr ? reinterpret_cast<const uchar4*>(reinterpret_cast<const uint8_t*>(cv->d_src) + (size_t)k * img_h * cv->src_step_px)
                            : cv->d_src + (size_t)k * img_h * cv->src_step_px;
        d.src_row0 = k * img_h;
        if (!d.keep) continue;
        d.chip = cv->d_chips + coff[k] / 4; d.mask = cv->d_masks + moff[k];
    }
    // K5 stages the source footprint of a chip tile with TMA boxes of kFpBoxW x kFpBoxH pixels
    cv->tmap_src_ok = false;
    if (cv->src_bgr) {
        // packed BGR rows as a tensor of u32 words: 3 W / 4 words per row (W % 16 == 0), box = 132 words (176 px) x 16 rows
        if ((int64_t)n_images * img_h < ((int64_t)1 << 31) &&
            uavm_encode_tmap_2d(ctx, &cv->tmap_src, (int)CU_TENSOR_MAP_DATA_TYPE_UINT32, cv->d_src, (uint64_t)cv->src_step_px / 4, (uint64_t)n_images * img_h,
                                (uint64_t)cv->src_step_px, kFpPitchBgr / 4, kFpBoxH, 0) == UAVM_OK)
            cv->tmap_src_ok = true;
    } else if ((cv->src_step_px & 3) == 0 && (int64_t)n_images * img_h < ((int64_t)1 << 31) &&
        uavm_encode_tmap_2d(ctx, &cv->tmap_src, (int)CU_TENSOR_MAP_DATA_TYPE_UINT32, cv->d_src, (uint64_t)cv->src_step_px, (uint64_t)n_images * img_h,
                            (uint64_t)cv->src_step_px * 4, kFpBoxW, kFpBoxH, 0) == UAVM_OK)
        cv->tmap_src_ok = true;
    cv->rect_x0 = 0; cv->rect_y0 = 0; cv->rect_x1 = cv->layout.canvas_w; cv->rect_y1 = cv->layout.canvas_h;
    cv->need_base.resize((size_t)n_images * 4);
    for (int k = 0; k < n_images; k++) { cv->need_base[4 * k] = cv->desc[k].need_x0; cv->need_base[4 * k + 1] = cv->desc[k].need_y0; cv->need_base[4 * k + 2] = cv->desc[k].need_x1; cv->need_base[4 * k + 3] = cv->desc[k].need_y1; }
    rc = uavm_canvas_upload_desc(ctx, cv);
    if (rc != UAVM_OK) { uavm_canvas_destroy(ctx, cv); return rc; }
    *out = cv;
    return UAVM_OK;
}

// Multi-GPU canvas sharding: this context produces only the canvas rectangle [x0, x1) x [y0, y1) (even edges; x1 / y1 may be
// the canvas size).  K7 computes the pyramid rectangles the output depends on (blend_plan.h), so the result is bit-identical to
// the unsharded blend without any halo parameter.  Chips that cannot contribute are deactivated: they are neither warped nor
// masked nor fed; active chips are warped only where K7 can read them.  The full canvas rectangle restores the default.
extern "C" int uavm_canvas_set_rect(uavm_ctx* ctx, uavm_canvas* cv, int x0, int y0, int x1, int y1)
{
    if (!ctx || !cv) return UAVM_EINVAL;
    const int cw = cv->layout.canvas_w, ch = cv->layout.canvas_h;
    if (x0 < 0 || y0 < 0 || x1 > cw || y1 > ch || x0 >= x1 || y0 >= y1 || (x0 & 1) || (y0 & 1) || ((x1 & 1) && x1 != cw) || ((y1 & 1) && y1 != ch)) {
        UAVM_SET_ERR(ctx, "set_rect: edges must be even and inside the canvas"); return UAVM_EINVAL;
    }
    cv->rect_x0 = x0; cv->rect_y0 = y0; cv->rect_x1 = x1; cv->rect_y1 = y1;
    cv->sharded = !(x0 == 0 && y0 == 0 && x1 == cw && y1 == ch);
    // What can K7 read of a chip?  Plan the chip with its whole box as the mask box, for the deepest pyramid the blender
    // supports at this canvas size: pyramid dependencies grow with the band count, so this covers every num_bands <= 5
    // (the reference's constant, M/MosaicWithoutPos.cpp:2179) — uavm_canvas_blend rejects a sharded blend with more bands.
    using namespace uavm_plan;
    const double max_len = (double)(cw > ch ? cw : ch);
    int nb = (int)ceil(log(max_len) / log(2.0)); if (nb > 5) nb = 5;
    const int M = uavm_canvas::kShardMargin;
    const IRect seam{x0 - M > 0 ? x0 - M : 0, y0 - M > 0 ? y0 - M : 0, x1 + M < cw ? x1 + M : cw, y1 + M < ch ? y1 + M : ch};   // where K6 writes masks
    cv->max_chip_w = 0; cv->max_chip_h = 0;
    for (int k = 0; k < cv->n; k++) {
        ChipDesc& d = cv->desc[k];
        const uavm_chip_layout& c = cv->chips[k];
        d.keep = c.keep ? 1 : 0;
        d.need_x0 = 0; d.need_y0 = 0; d.need_x1 = c.chip_w; d.need_y1 = c.chip_h;
        if (c.keep && cv->sharded) {
            IRect need = make_empty();
            for (int b = 0; b <= nb; b++) {
                const CanvasPlan P = plan_canvas(cw, ch, b, IRect{x0, y0, x1, y1});
                const IRect a0 = shifted(isect(IRect{c.beg_x, c.beg_y, c.beg_x + c.chip_w, c.beg_y + c.chip_h}, seam), -c.beg_x, -c.beg_y);
                const ChipPlan cp = plan_chip(c.beg_x, c.beg_y, c.chip_w, c.chip_h, a0, P);
                if (!cp.active) continue;
                int lx, hx, ly, hy;
                reflect_range(cp.C[0].x0 - cp.roi.left, cp.C[0].x1 - cp.roi.left, c.chip_w, lx, hx);
                reflect_range(cp.C[0].y0 - cp.roi.top, cp.C[0].y1 - cp.roi.top, c.chip_h, ly, hy);
                if (hx > lx && hy > ly) need = hull(need, IRect{lx, ly, hx, hy});
            }
            if (is_empty(need)) d.keep = 0;
            else { d.need_x0 = need.x0; d.need_y0 = need.y0; d.need_x1 = need.x1; d.need_y1 = need.y1; }
        }
        if (d.keep) { if (c.chip_w > cv->max_chip_w) cv->max_chip_w = c.chip_w; if (c.chip_h > cv->max_chip_h) cv->max_chip_h = c.chip_h; }
    }
    cv->lines_dirty = true; cv->warped = false; cv->seamed = false; cv->blended = false; cv->mask_plane_valid = false; cv->own_bbox_valid = false;
    for (int k = 0; k < cv->n; k++) { cv->need_base[4 * k] = cv->desc[k].need_x0; cv->need_base[4 * k + 1] = cv->desc[k].need_y0; cv->need_base[4 * k + 2] = cv->desc[k].need_x1; cv->need_base[4 * k + 3] = cv->desc[k].need_y1; }
    cv->need_narrowed = false;
    return uavm_canvas_upload_desc(ctx, cv);
}
// horizontal band [y0, y1) of the canvas.  `halo` is accepted for source compatibility and ignored: K7 derives what a
// rectangle depends on from the band count instead of trusting a caller-supplied halo.
extern "C" int uavm_canvas_set_band(uavm_ctx* ctx, uavm_canvas* cv, int y0, int y1, int halo)
{
    (void)halo;
    if (!ctx || !cv) return UAVM_EINVAL;
    return uavm_canvas_set_rect(ctx, cv, 0, y0, cv->layout.canvas_w, y1);
}
extern "C" int uavm_canvas_is_active(uavm_canvas* cv, int image)
{
    if (!cv || image < 0 || image >= cv->n) return 0;
    return cv->desc[image].keep;
}

extern "C" void uavm_canvas_destroy(uavm_ctx* ctx, uavm_canvas* cv)
{
    if (!cv) return;
    if (ctx) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream); }
    for (int s = 0; s < uavm_canvas::kStageSlots; s++) {
        cudaFree(cv->d_stage[s]);
        if (cv->ev_copied[s]) cudaEventDestroy(cv->ev_copied[s]);
        if (cv->ev_free[s]) cudaEventDestroy(cv->ev_free[s]);
    }
    for (int e = 0; e < uavm_canvas::kWarpEvents; e++) if (cv->ev_warp[e]) cudaEventDestroy(cv->ev_warp[e]);
    uavm_blend_free(cv);
    cudaFree(cv->d_src); cudaFree(cv->d_chips); cudaFree(cv->d_masks); cudaFree(cv->d_dist_max); cudaFree(cv->d_k6); cudaFree(cv->d_mask_ptr); cudaFree(cv->d_mask_step); cudaFree(cv->d_own_bbox);
    if (cv->bound_dist) uavm_dist_forget_canvas(cv->bound_dist, cv);
    cudaFree(cv->d_desc); cudaFree(cv->d_result); cudaFree(cv->d_result_mask);
    delete cv;
}

// bytes per source pixel in HBM: 3 = the caller's BGR frames are kept as they are (width % 16 == 0), 4 = BGRA pool + conversion pass
extern "C" int uavm_canvas_source_layout(const uavm_canvas* cv) { return cv ? (cv->src_bgr ? 3 : 4) : 0; }

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
    UAVM_CUDA(ctx, cudaSetDevice(ctx->device));
    if (cv->src_bgr) {
        // the pool keeps the caller's BGR bytes: one copy straight into the frame's slot, no conversion.  Host frames travel on
        // the copy stream (the compute stream only waits for THIS frame); the copy first waits for the last warp that read the slot.
        uint8_t* dst8 = reinterpret_cast<uint8_t*>(cv->d_src) + (size_t)image * cv->img_h * cv->src_step_px;
        cudaStream_t cs = is_device ? ctx->stream : ctx->copy_stream;
        if (!is_device && cv->last_warp_ev[image] >= 0) UAVM_CUDA(ctx, cudaStreamWaitEvent(cs, cv->ev_warp[cv->last_warp_ev[image]], 0));
        const cudaMemcpyKind kind = is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
        if (step == cv->src_step_px) UAVM_CUDA(ctx, cudaMemcpyAsync(dst8, bgr, (size_t)step * cv->img_h, kind, cs));
        else UAVM_CUDA(ctx, cudaMemcpy2DAsync(dst8, (size_t)cv->src_step_px, bgr, (size_t)step, (size_t)cv->img_w * 3, cv->img_h, kind, cs));
        if (!is_device) {
            const int slot = cv->stage_next; cv->stage_next = (slot + 1) % uavm_canvas::kStageSlots;
            UAVM_CUDA(ctx, cudaEventRecord(cv->ev_copied[slot], cs));
            UAVM_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, cv->ev_copied[slot], 0));
        }
        return UAVM_OK;
    }
    const uint8_t* src = bgr; int sstep = step;
    int slot = -1;
    if (!is_device) {
        // host frame: PCIe copy on the copy stream into a ring slot; the compute stream only waits for THIS frame
        slot = cv->stage_next; cv->stage_next = (slot + 1) % uavm_canvas::kStageSlots;
        UAVM_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, cv->ev_free[slot], 0));          // previous user of the slot has been converted
        if (step == cv->stage_pitch)                     // contiguous frame: one linear DMA (the 2-D path is slower over PCIe)
            UAVM_CUDA(ctx, cudaMemcpyAsync(cv->d_stage[slot], bgr, (size_t)step * cv->img_h, cudaMemcpyHostToDevice, ctx->copy_stream));
        else
            UAVM_CUDA(ctx, cudaMemcpy2DAsync(cv->d_stage[slot], (size_t)cv->stage_pitch, bgr, (size_t)step, (size_t)cv->img_w * 3, cv->img_h,
                                             cudaMemcpyHostToDevice, ctx->copy_stream));
        UAVM_CUDA(ctx, cudaEventRecord(cv->ev_copied[slot], ctx->copy_stream));
        UAVM_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, cv->ev_copied[slot], 0));
        src = cv->d_stage[slot]; sstep = cv->stage_pitch;
    }
    uchar4* dst = cv->d_src + (size_t)image * cv->img_h * cv->src_step_px;
    if ((sstep % 4) == 0 && (reinterpret_cast<uintptr_t>(src) % 4) == 0) {
        dim3 grid(((cv->img_w + 3) / 4 + 127) / 128, cv->img_h);
        k5_bgr_to_bgra_x4<<<grid, 128, 0, ctx->stream>>>(src, cv->img_w, sstep, dst, cv->src_step_px);
    } else {
        dim3 grid((cv->img_w + 255) / 256, cv->img_h);
        k5_bgr_to_bgra<<<grid, 256, 0, ctx->stream>>>(src, cv->img_w, cv->img_h, sstep, dst, cv->src_step_px);
    }
    UAVM_CHECK_LAUNCH(ctx);
    if (slot >= 0) UAVM_CUDA(ctx, cudaEventRecord(cv->ev_free[slot], ctx->stream));
    return UAVM_OK;
}

// images [first, first + count): lets a caller that streams frames in warp each group as soon as it has been set
extern "C" int uavm_canvas_warp_range(uavm_ctx* ctx, uavm_canvas* cv, int first, int count)
{
    if (!ctx || !cv || first < 0 || count < 0 || first + count > cv->n) return UAVM_EINVAL;
    if (cv->max_chip_w <= 0 || count == 0) return UAVM_OK;
    UAVM_CUDA(ctx, cudaSetDevice(ctx->device));
    if (cv->need_narrowed && !cv->in_warp_for_blend) {         // a plain warp after uavm_canvas_warp_for_blend produces whole chips again
        for (int k = 0; k < cv->n; k++) { cv->desc[k].need_x0 = cv->need_base[4 * k]; cv->desc[k].need_y0 = cv->need_base[4 * k + 1]; cv->desc[k].need_x1 = cv->need_base[4 * k + 2]; cv->desc[k].need_y1 = cv->need_base[4 * k + 3]; }
        cv->need_narrowed = false;
        int rcu = uavm_canvas_upload_desc(ctx, cv); if (rcu != UAVM_OK) return rcu;
    }
    dim3 grid((cv->max_chip_w + kWarpTileW - 1) / kWarpTileW, (cv->max_chip_h + kWarpTileH - 1) / kWarpTileH, count);
    dim3 block(32, kWarpsY);
    bool any_affine = false, any_proj = false;
    for (int k = first; k < first + count; k++)
        if (cv->desc[k].keep) { if (cv->desc[k].affine) any_affine = true; else any_proj = true; }
    if (any_affine) {
        static const bool scalar = getenv("UAVM_K5_SCALAR") != nullptr;       // A/B switch for profiling; both are bit-exact
        // the packed kernel derives tap offsets from 23-bit mantissas and 32-bit pixel offsets
        if (!scalar && cv->img_w < (1 << 22) && cv->img_h < (1 << 22) && (int64_t)cv->img_h * cv->src_step_px * (cv->src_bgr ? 1 : 4) < ((int64_t)1 << 32)) {
            static const int env_boxes = getenv("UAVM_K5_BOXES") ? atoi(getenv("UAVM_K5_BOXES")) : kFpBoxes;   // 0: direct tap loads (A/B)
            const int boxes = cv->tmap_src_ok ? env_boxes : 0;
            if (!ctx->k5_attr_set) {                      // function attributes are per device: once per context
                cudaFuncSetAttribute(k5_warp_affine_x2<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
                cudaFuncSetAttribute(k5_warp_affine_x2<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
                ctx->k5_attr_set = true;
            }
            if (cv->src_bgr)
                k5_warp_affine_x2<true><<<grid, block, boxes * kFpBoxBytesBgr, ctx->stream>>>(
                    cv->tmap_src, cv->d_desc + first, cv->img_w, cv->img_h, cv->src_step_px, cv->layout.dgx, cv->layout.dgy, 0x4B000000u,
                    (float)(cv->img_w - 1), (float)(cv->img_h - 1), 1.0f, 1u, 0x4B000000ull * ((unsigned long long)cv->src_step_px + 3ull), boxes);
            else
                k5_warp_affine_x2<false><<<grid, block, boxes * kFpBoxBytes, ctx->stream>>>(
                    cv->tmap_src, cv->d_desc + first, cv->img_w, cv->img_h, cv->src_step_px, cv->layout.dgx, cv->layout.dgy, 0x4B000000u,
                    (float)(cv->img_w - 1), (float)(cv->img_h - 1), 1.0f, 1u, 0x4B000000ull * (4ull * (unsigned long long)cv->src_step_px + 4ull), boxes);
        }
        else if (cv->src_bgr)
            k5_warp_chips<true, true><<<grid, block, 0, ctx->stream>>>(cv->d_desc + first, cv->img_w, cv->img_h, cv->src_step_px, cv->layout.dgx, cv->layout.dgy, 0x4B000000u,
                                                                       (float)(cv->img_w - 1), (float)(cv->img_h - 1));
        else
            k5_warp_chips<true, false><<<grid, block, 0, ctx->stream>>>(cv->d_desc + first, cv->img_w, cv->img_h, cv->src_step_px, cv->layout.dgx, cv->layout.dgy, 0x4B000000u,
                                                                        (float)(cv->img_w - 1), (float)(cv->img_h - 1));
        UAVM_CHECK_LAUNCH(ctx);
    }
    if (any_proj) {
        if (cv->src_bgr)
            k5_warp_chips<false, true><<<grid, block, 0, ctx->stream>>>(cv->d_desc + first, cv->img_w, cv->img_h, cv->src_step_px, cv->layout.dgx, cv->layout.dgy, 0x4B000000u,
                                                                        (float)(cv->img_w - 1), (float)(cv->img_h - 1));
        else
            k5_warp_chips<false, false><<<grid, block, 0, ctx->stream>>>(cv->d_desc + first, cv->img_w, cv->img_h, cv->src_step_px, cv->layout.dgx, cv->layout.dgy, 0x4B000000u,
                                                                         (float)(cv->img_w - 1), (float)(cv->img_h - 1));
        UAVM_CHECK_LAUNCH(ctx);
    }
    if (cv->src_bgr) {                                    // uploads into these frames' slots must wait for this launch
        const int e = cv->ev_warp_next; cv->ev_warp_next = (e + 1) % uavm_canvas::kWarpEvents;
        UAVM_CUDA(ctx, cudaEventRecord(cv->ev_warp[e], ctx->stream));
        for (int k = first; k < first + count; k++) cv->last_warp_ev[k] = e;
    }
    cv->warped = true; cv->seamed = false; cv->mask_plane_valid = false;
    return UAVM_OK;
}

extern "C" int uavm_canvas_warp(uavm_ctx* ctx, uavm_canvas* cv)
{
    if (!ctx || !cv) return UAVM_EINVAL;
    return uavm_canvas_warp_range(ctx, cv, 0, cv->n);
}

// K5 for a mosaic: warp only the chip pixels the blend can read.  With seam masks a chip contributes where it owns canvas pixels
// (plus the support of the blender's pyramids), typically a third of its area; uavm_canvas_seam_masks does not need the chips
// (validity is recomputed from the coordinates), so the pipeline runs K6 first, takes the bounding box of what each chip owns
// and warps the rectangle a blend of up to 5 bands reads (blend_plan.h).  The mosaic is bit-identical; uavm_canvas_get_chip
// afterwards returns a partly filled chip.  A plain uavm_canvas_warp restores whole chips.
extern "C" int uavm_canvas_warp_for_blend(uavm_ctx* ctx, uavm_canvas* cv)
{
    if (!ctx || !cv) return UAVM_EINVAL;
    if (!cv->seamed) { UAVM_SET_ERR(ctx, "warp_for_blend needs the seam masks (uavm_canvas_seam_masks) first"); return UAVM_EINVAL; }
    UAVM_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!cv->own_bbox_valid) {
        cv->own_bbox.resize((size_t)cv->n * 4);
        UAVM_CUDA(ctx, cudaMemcpyAsync(cv->own_bbox.data(), cv->d_own_bbox, (size_t)cv->n * 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
        UAVM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        cv->own_bbox_valid = true;
    }
    using namespace uavm_plan;
    const int cw = cv->layout.canvas_w, ch = cv->layout.canvas_h;
    const double max_len = (double)(cw > ch ? cw : ch);
    int nbmax = (int)ceil(log(max_len) / log(2.0)); if (nbmax > 5) nbmax = 5;
    const IRect out = cv->sharded ? IRect{cv->rect_x0, cv->rect_y0, cv->rect_x1, cv->rect_y1} : IRect{0, 0, cw, ch};
    for (int k = 0; k < cv->n; k++) {
        ChipDesc& d = cv->desc[k];
        if (!d.keep) continue;
        const int32_t* b = &cv->own_bbox[(size_t)k * 4];
        IRect need = make_empty();
        if (!(b[2] < b[0] || b[3] < b[1])) {
            const IRect a0{b[0], b[1], b[2] + 1, b[3] + 1};
            for (int nb = 0; nb <= nbmax; nb++) {
                const CanvasPlan P = plan_canvas(cw, ch, nb, out);
                const ChipPlan cp = plan_chip(d.beg_x, d.beg_y, d.chip_w, d.chip_h, a0, P);
                if (!cp.active) continue;
                int lx, hx, ly, hy;
                reflect_range(cp.C[0].x0 - cp.roi.left, cp.C[0].x1 - cp.roi.left, d.chip_w, lx, hx);
                reflect_range(cp.C[0].y0 - cp.roi.top, cp.C[0].y1 - cp.roi.top, d.chip_h, ly, hy);
                if (hx > lx && hy > ly) need = hull(need, IRect{lx, ly, hx, hy});
            }
        }
        // never wider than what create / set_rect allowed
        const IRect base{cv->need_base[4 * k], cv->need_base[4 * k + 1], cv->need_base[4 * k + 2], cv->need_base[4 * k + 3]};
        need = isect(need, base);
        d.need_x0 = need.x0; d.need_y0 = need.y0; d.need_x1 = need.x1; d.need_y1 = need.y1;      // empty: nothing of this chip is read
    }
    cv->need_narrowed = true;
    int rc = uavm_canvas_upload_desc(ctx, cv);
    if (rc != UAVM_OK) return rc;
    cv->in_warp_for_blend = true;
    rc = uavm_canvas_warp_range(ctx, cv, 0, cv->n);
    cv->in_warp_for_blend = false;
    cv->seamed = true; cv->mask_plane_valid = true;       // warp_range reset them: the seam masks are still the current ones
    return rc;
}

// expands the validity masks (chip alpha) into the u8 mask plane unless the plane already holds them or K6's seam masks
int uavm_canvas_mask_plane(uavm_ctx* ctx, uavm_canvas* cv)
{
    if (cv->seamed || cv->mask_plane_valid || cv->max_chip_w <= 0) return UAVM_OK;
    dim3 grid(((cv->max_chip_w + 3) / 4 + 255) / 256, cv->max_chip_h, cv->n);
    k5_alpha_to_mask<<<grid, 256, 0, ctx->stream>>>(cv->d_desc);
    UAVM_CHECK_LAUNCH(ctx);
    cv->mask_plane_valid = true;
    return UAVM_OK;
}

extern "C" int uavm_canvas_get_chip(uavm_ctx* ctx, uavm_canvas* cv, int image, uint8_t* chip_bgr, int chip_step, uint8_t* mask, int mask_step)
{
    if (!ctx || !cv || image < 0 || image >= cv->n) return UAVM_EINVAL;
    const ChipDesc& d = cv->desc[image];
    if (!d.keep) return UAVM_EINVAL;
    uint8_t* tmp = nullptr;
    if (chip_bgr) {
        if (chip_step < d.chip_w * 3) return UAVM_EINVAL;
        UAVM_CUDA(ctx, cudaMalloc(&tmp, (size_t)d.chip_w * d.chip_h * 3));
        dim3 grid((d.chip_w + 255) / 256, d.chip_h);
        k5_chip_to_bgr<<<grid, 256, 0, ctx->stream>>>(d.chip, d.chip_step, d.chip_w, d.chip_h, tmp);
        if (cudaGetLastError() != cudaSuccess) { cudaFree(tmp); return UAVM_EFAIL; }
        ctx->launches++;
        if (cudaMemcpy2DAsync(chip_bgr, (size_t)chip_step, tmp, (size_t)d.chip_w * 3, (size_t)d.chip_w * 3, d.chip_h, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) {
            cudaFree(tmp); return UAVM_EFAIL;
        }
    }
    if (mask) {
        if (mask_step < d.chip_w) { cudaFree(tmp); return UAVM_EINVAL; }
        int rc = uavm_canvas_mask_plane(ctx, cv);
        if (rc != UAVM_OK) { cudaFree(tmp); return rc; }
        if (cudaMemcpy2DAsync(mask, (size_t)mask_step, d.mask, (size_t)d.mask_step, (size_t)d.chip_w, d.chip_h, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) {
            cudaFree(tmp); return UAVM_EFAIL;
        }
    }
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    cudaFree(tmp);
    if (e != cudaSuccess) { UAVM_SET_ERR(ctx, "get_chip: %s", cudaGetErrorString(e)); return UAVM_EFAIL; }
    return UAVM_OK;
}
