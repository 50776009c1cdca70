// sift.cu — f1: SIFT detection + description on the GPU.  Replaces SiftExtraction_Thread's
// `SIFT(2000, 3, 0.01, 20)` detect + compute (M/MosaicWithoutPos.cpp:4852-4872; OpenCV nonfree 2.4.0, third party, binary
// only — CV/nonfree/features2d.hpp:58-100).  The algorithm restated here is OpenCV's published one (features2d/src/sift*.cpp,
// unchanged in structure since 2.4): gray -> float -> 2x linear upsample -> Gaussian scale space (nOctaveLayers + 3 images per
// octave, incremental sigmas) -> DoG -> 26-neighbour extrema -> sub-pixel / sub-scale refinement (<= 5 Newton steps), contrast
// and edge tests -> 36-bin orientation histograms (peaks >= 0.8 max, parabolic interpolation) -> KeyPointsFilter
// (removeDuplicatedSorted, retainBest) -> 4 x 4 x 8 descriptors (trilinear votes, Gaussian window, normalise, clip 0.2,
// renormalise, x 512, saturate to u8).
// Parity (proxy: cv2 4.13 SIFT_create, tests/test_gpu_sift.py): the u8 -> gray conversion and the 2x upsample are bit-exact;
// the Gaussian blurs agree to ~2e-7 relative (OpenCV's float filter runs through IPP / SIMD with another summation order), so
// keypoints agree to sub-pixel noise except for decisions that sit on a threshold, and descriptor bytes differ by at most a
// few LSB.  SIFT parity with the reference's 2.4.0 binary is UNPINNED (no keypoints ship with the reference).
//
// GPU organisation: one image at a time, the whole scale space resident (a 4000 x 3000 frame: 8000 x 6000 base, 2.8 GB);
// separable blurs with the taps in constant memory; one thread per DoG pixel for the extremum test, the rare survivors are
// refined in place and appended to a candidate list; one warp per candidate for the orientation histogram; one CTA per kept
// keypoint for the descriptor, votes accumulated as 64-bit fixed point in shared memory so the sums do not depend on the
// order the threads arrive in (deterministic output).
#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include "internal.h"

namespace {

constexpr int kMaxOctaves = 16, kMaxLayers = 8, kMaxTaps = 64;
constexpr int kImgBorder = 5, kMaxInterp = 5, kOriBins = 36, kDescW = 4, kDescBins = 8;
constexpr float kOriSigFctr = 1.5f, kOriRadius = 4.5f, kOriPeakRatio = 0.8f, kDescSclFctr = 3.0f, kDescMagThr = 0.2f, kIntDescrFctr = 512.0f;

__constant__ float c_taps[kMaxLayers + 3][kMaxTaps];      // [blur index][0 = centre .. radius]
__constant__ int c_radius[kMaxLayers + 3];

struct Cand { float x, y, xi, size, response; int32_t o, layer, r, c; };                   // x, y in octave pixels (c + xc, r + xr)
struct Kp { float x, y, size, angle, response; int32_t octave; };

__device__ __forceinline__ int reflect101(int p, int n) { if (n == 1) return 0; while (p < 0 || p >= n) p = p < 0 ? -p : 2 * n - 2 - p; return p; }

// BGR u8 -> gray (cvtColor BGR2GRAY: (B 3735 + G 19235 + R 9798 + 2^14) >> 15) -> float -> 2x INTER_LINEAR upsample
__global__ void __launch_bounds__(256) ks_init(const uint8_t* __restrict__ bgr, int step, int w, int h, float* __restrict__ out)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= 2 * w || y >= 2 * h) return;
    auto coef = [](int d, int n, int& s0, int& s1, float& a) {
        const float f = ((float)d + 0.5f) * 0.5f - 0.5f;
        int s = (int)floorf(f); a = f - (float)s;
        if (s < 0) { s = 0; a = 0.0f; }
        if (s >= n - 1) { s = n - 1; a = 0.0f; }
        s0 = s; s1 = min(s + 1, n - 1);
    };
    int x0, x1, y0, y1; float ax, ay;
    coef(x, w, x0, x1, ax); coef(y, h, y0, y1, ay);
    auto gray = [&](int yy, int xx) {
        const uint8_t* p = bgr + (size_t)yy * step + 3 * xx;
        return (float)((p[0] * 3735 + p[1] * 19235 + p[2] * 9798 + (1 << 14)) >> 15);
    };
    const float h0 = gray(y0, x0) * (1.0f - ax) + gray(y0, x1) * ax;
    const float h1 = gray(y1, x0) * (1.0f - ax) + gray(y1, x1) * ax;
    out[(size_t)y * (2 * w) + x] = h0 * (1.0f - ay) + h1 * ay;
}

// separable Gaussian, BORDER_REFLECT_101; symmetric form k0 x0 + sum_k k_k (x_-k + x_+k)
template <bool ROWS>
__global__ void __launch_bounds__(256) ks_blur(const float* __restrict__ src, float* __restrict__ dst, int w, int h, int ki)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w || y >= h) return;
    const int r = c_radius[ki];
    const float* k = c_taps[ki];
    float s;
    if (ROWS) {
        const float* row = src + (size_t)y * w;
        s = k[0] * row[x];
        if (x >= r && x + r < w) { for (int t = 1; t <= r; t++) s = s + k[t] * (row[x - t] + row[x + t]); }
        else { for (int t = 1; t <= r; t++) s = s + k[t] * (row[reflect101(x - t, w)] + row[reflect101(x + t, w)]); }
    } else {
        s = k[0] * src[(size_t)y * w + x];
        if (y >= r && y + r < h) { for (int t = 1; t <= r; t++) s = s + k[t] * (src[(size_t)(y - t) * w + x] + src[(size_t)(y + t) * w + x]); }
        else { for (int t = 1; t <= r; t++) s = s + k[t] * (src[(size_t)reflect101(y - t, h) * w + x] + src[(size_t)reflect101(y + t, h) * w + x]); }
    }
    dst[(size_t)y * w + x] = s;
}

__global__ void __launch_bounds__(256) ks_down2(const float* __restrict__ src, int sw, float* __restrict__ dst, int w, int h)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x < w && y < h) dst[(size_t)y * w + x] = src[(size_t)(2 * y) * sw + 2 * x];          // INTER_NEAREST to exactly half the size
}

__global__ void __launch_bounds__(256) ks_sub(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ d, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) d[i] = a[i] - b[i];
}

struct OctaveDog { const float* dog[kMaxLayers + 2]; int w, h; };

// adjustLocalExtrema: Newton steps on the 3-D quadratic, then the contrast and edge tests
__device__ bool adjust_extremum(const OctaveDog& D, int nl, int& layer, int& r, int& c, float contrast_thr, float edge_thr, float sigma, int octv, Cand& out)
{
    const float img_scale = 1.0f / 255.0f, deriv_scale = img_scale * 0.5f, second_deriv_scale = img_scale, cross_deriv_scale = img_scale * 0.25f;
    float xi = 0, xr = 0, xc = 0;
    int i = 0;
    for (; i < kMaxInterp; i++) {
        const float* img = D.dog[layer]; const float* prv = D.dog[layer - 1]; const float* nxt = D.dog[layer + 1];
        const size_t o = (size_t)r * D.w + c;
        const float dDx = (img[o + 1] - img[o - 1]) * deriv_scale, dDy = (img[o + D.w] - img[o - D.w]) * deriv_scale, dDs = (nxt[o] - prv[o]) * deriv_scale;
        const float v2 = img[o] * 2.0f;
        const float dxx = (img[o + 1] + img[o - 1] - v2) * second_deriv_scale;
        const float dyy = (img[o + D.w] + img[o - D.w] - v2) * second_deriv_scale;
        const float dss = (nxt[o] + prv[o] - v2) * second_deriv_scale;
        const float dxy = (img[o + D.w + 1] - img[o + D.w - 1] - img[o - D.w + 1] + img[o - D.w - 1]) * cross_deriv_scale;
        const float dxs = (nxt[o + 1] - nxt[o - 1] - prv[o + 1] + prv[o - 1]) * cross_deriv_scale;
        const float dys = (nxt[o + D.w] - nxt[o - D.w] - prv[o + D.w] + prv[o - D.w]) * cross_deriv_scale;
        // X = H^-1 dD (Matx33f::solve, closed form), offsets = -X
        const float a00 = dxx, a01 = dxy, a02 = dxs, a11 = dyy, a12 = dys, a22 = dss;
        float det = a00 * (a11 * a22 - a12 * a12) - a01 * (a01 * a22 - a12 * a02) + a02 * (a01 * a12 - a11 * a02);
        float X0 = 0, X1 = 0, X2 = 0;
        if (det != 0.0f) {
            det = 1.0f / det;
            X0 = det * (dDx * (a11 * a22 - a12 * a12) - a01 * (dDy * a22 - a12 * dDs) + a02 * (dDy * a12 - a11 * dDs));
            X1 = det * (a00 * (dDy * a22 - a12 * dDs) - dDx * (a01 * a22 - a12 * a02) + a02 * (a01 * dDs - dDy * a02));
            X2 = det * (a00 * (a11 * dDs - dDy * a12) - a01 * (a01 * dDs - dDy * a02) + dDx * (a01 * a12 - a11 * a02));
        }
        xi = -X2; xr = -X1; xc = -X0;
        if (fabsf(xi) < 0.5f && fabsf(xr) < 0.5f && fabsf(xc) < 0.5f) break;
        if (fabsf(xi) > (float)(INT_MAX / 3) || fabsf(xr) > (float)(INT_MAX / 3) || fabsf(xc) > (float)(INT_MAX / 3)) return false;
        c += __float2int_rn(xc); r += __float2int_rn(xr); layer += __float2int_rn(xi);
        if (layer < 1 || layer > nl || c < kImgBorder || c >= D.w - kImgBorder || r < kImgBorder || r >= D.h - kImgBorder) return false;
    }
    if (i >= kMaxInterp) return false;
    {
        const float* img = D.dog[layer]; const float* prv = D.dog[layer - 1]; const float* nxt = D.dog[layer + 1];
        const size_t o = (size_t)r * D.w + c;
        const float dDx = (img[o + 1] - img[o - 1]) * deriv_scale, dDy = (img[o + D.w] - img[o - D.w]) * deriv_scale, dDs = (nxt[o] - prv[o]) * deriv_scale;
        const float t = dDx * xc + dDy * xr + dDs * xi;
        const float contr = img[o] * img_scale + t * 0.5f;
        if (fabsf(contr) * (float)nl < contrast_thr) return false;
        const float v2 = img[o] * 2.0f;
        const float dxx = (img[o + 1] + img[o - 1] - v2) * second_deriv_scale;
        const float dyy = (img[o + D.w] + img[o - D.w] - v2) * second_deriv_scale;
        const float dxy = (img[o + D.w + 1] - img[o + D.w - 1] - img[o - D.w + 1] + img[o - D.w - 1]) * cross_deriv_scale;
        const float tr = dxx + dyy, det = dxx * dyy - dxy * dxy;
        if (det <= 0.0f || tr * tr * edge_thr >= (edge_thr + 1.0f) * (edge_thr + 1.0f) * det) return false;
        out.x = (float)c + xc; out.y = (float)r + xr; out.xi = xi; out.o = octv; out.layer = layer; out.r = r; out.c = c;
        out.size = sigma * powf(2.0f, ((float)layer + xi) / (float)nl) * (float)(1 << octv) * 2.0f;
        out.response = fabsf(contr);
    }
    return true;
}

__global__ void __launch_bounds__(256)
ks_extrema(OctaveDog D, int nl, int octv, int threshold, float contrast_thr, float edge_thr, float sigma, Cand* __restrict__ cand, int* __restrict__ n_cand, int cap)
{
    const int c = kImgBorder + blockIdx.x * blockDim.x + threadIdx.x, r = kImgBorder + blockIdx.y, layer0 = 1 + blockIdx.z;
    if (c >= D.w - kImgBorder || r >= D.h - kImgBorder) return;
    const float* img = D.dog[layer0]; const float* prv = D.dog[layer0 - 1]; const float* nxt = D.dog[layer0 + 1];
    const size_t o = (size_t)r * D.w + c;
    const float val = img[o];
    if (!(fabsf(val) > (float)threshold)) return;
    bool ext = true;
    if (val > 0) {
#pragma unroll
        for (int dy = -1; dy <= 1 && ext; dy++)
#pragma unroll
            for (int dx = -1; dx <= 1; dx++) {
                const size_t q = o + (ptrdiff_t)dy * D.w + dx;
                if (!(val >= img[q] && val >= prv[q] && val >= nxt[q])) { ext = false; break; }
            }
    } else {
#pragma unroll
        for (int dy = -1; dy <= 1 && ext; dy++)
#pragma unroll
            for (int dx = -1; dx <= 1; dx++) {
                const size_t q = o + (ptrdiff_t)dy * D.w + dx;
                if (!(val <= img[q] && val <= prv[q] && val <= nxt[q])) { ext = false; break; }
            }
    }
    if (!ext) return;
    int layer = layer0, rr = r, cc = c;
    Cand k;
    if (!adjust_extremum(D, nl, layer, rr, cc, contrast_thr, edge_thr, sigma, octv, k)) return;
    const int slot = atomicAdd(n_cand, 1);
    if (slot < cap) cand[slot] = k;
}

// fastAtan2 of OpenCV (degrees, polynomial, max error ~0.3 deg)
__device__ __forceinline__ float fast_atan2_deg(float y, float x)
{
    const float p1 = 0.9997878412794807f * 57.29577951308232f, p3 = -0.3258083974640975f * 57.29577951308232f;
    const float p5 = 0.1555786518463281f * 57.29577951308232f, p7 = -0.04432655554792128f * 57.29577951308232f;
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) { c = ay / (ax + 2.220446049250313e-16f); c2 = c * c; a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c; }
    else { c = ax / (ay + 2.220446049250313e-16f); c2 = c * c; a = 90.0f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c; }
    if (x < 0) a = 180.0f - a;
    if (y < 0) a = 360.0f - a;
    return a;
}

struct OctaveGauss { const float* g[kMaxLayers + 3]; int w, h; };

// calcOrientationHist + peak extraction: one warp per candidate
__global__ void __launch_bounds__(256)
ks_orient(const Cand* __restrict__ cand, int n_cand, const OctaveGauss* __restrict__ octs, Kp* __restrict__ kps, int* __restrict__ n_kp, int cap)
{
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (wid >= n_cand) return;
    const Cand k = cand[wid];
    const OctaveGauss& G = octs[k.o];
    const float* img = G.g[k.layer];
    const float scl_octv = k.size * 0.5f / (float)(1 << k.o);
    const int radius = __float2int_rn(kOriRadius * scl_octv);
    const float sigma = kOriSigFctr * scl_octv, expf_scale = -1.0f / (2.0f * sigma * sigma);
    float hist[kOriBins];
#pragma unroll
    for (int b = 0; b < kOriBins; b++) hist[b] = 0.0f;
    const int side = 2 * radius + 1;
    for (int s = lane; s < side * side; s += 32) {
        const int i = s / side - radius, j = s % side - radius;
        const int y = k.r + i, x = k.c + j;
        if (y <= 0 || y >= G.h - 1 || x <= 0 || x >= G.w - 1) continue;
        const float dx = img[(size_t)y * G.w + x + 1] - img[(size_t)y * G.w + x - 1];
        const float dy = img[(size_t)(y - 1) * G.w + x] - img[(size_t)(y + 1) * G.w + x];
        const float wgt = expf((float)(i * i + j * j) * expf_scale);
        const float ori = fast_atan2_deg(dy, dx), mag = sqrtf(dx * dx + dy * dy);
        int bin = __float2int_rn(((float)kOriBins / 360.0f) * ori);
        if (bin >= kOriBins) bin -= kOriBins;
        if (bin < 0) bin += kOriBins;
#pragma unroll
        for (int b = 0; b < kOriBins; b++) if (b == bin) hist[b] += wgt * mag;      // register array: no dynamic indexing
    }
    // fixed-order butterfly: every lane ends with the same (deterministic) sums
#pragma unroll
    for (int b = 0; b < kOriBins; b++)
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) hist[b] += __shfl_xor_sync(0xffffffffu, hist[b], s);
    if (lane != 0) return;
    float sm[kOriBins], maxval = 0.0f;
#pragma unroll
    for (int b = 0; b < kOriBins; b++) {
        const float m2 = hist[(b + kOriBins - 2) % kOriBins], m1 = hist[(b + kOriBins - 1) % kOriBins], p1 = hist[(b + 1) % kOriBins], p2 = hist[(b + 2) % kOriBins];
        sm[b] = (m2 + p2) * (1.0f / 16.0f) + (m1 + p1) * (4.0f / 16.0f) + hist[b] * (6.0f / 16.0f);
        maxval = fmaxf(maxval, sm[b]);
    }
    const float mag_thr = maxval * kOriPeakRatio;
#pragma unroll
    for (int j = 0; j < kOriBins; j++) {
        const int l = j > 0 ? j - 1 : kOriBins - 1, r2 = j < kOriBins - 1 ? j + 1 : 0;
        if (sm[j] > sm[l] && sm[j] > sm[r2] && sm[j] >= mag_thr) {
            float bin = (float)j + 0.5f * (sm[l] - sm[r2]) / (sm[l] - 2.0f * sm[j] + sm[r2]);
            bin = bin < 0 ? (float)kOriBins + bin : bin >= (float)kOriBins ? bin - (float)kOriBins : bin;
            float angle = 360.0f - (360.0f / (float)kOriBins) * bin;
            if (fabsf(angle - 360.0f) < 1.1920929e-07f) angle = 0.0f;
            const int slot = atomicAdd(n_kp, 1);
            if (slot < cap) {
                Kp o;
                o.x = k.x * (float)(1 << k.o); o.y = k.y * (float)(1 << k.o); o.size = k.size; o.angle = angle; o.response = k.response;
                o.octave = k.o + (k.layer << 8) + (__float2int_rn((k.xi + 0.5f) * 255.0f) << 16);
                kps[slot] = o;
            }
        }
    }
}

// calcSIFTDescriptor: one CTA per keypoint.  kp holds FINAL keypoints (coordinates of the input image, octave already shifted by
// firstOctave = -1); first_octave = -1.
__global__ void __launch_bounds__(128)
ks_describe(const Kp* __restrict__ kps, int n, const OctaveGauss* __restrict__ octs, int nl, float* __restrict__ desc)
{
    constexpr int d = kDescW, nb = kDescBins, HL = (d + 2) * (d + 2) * (nb + 2);
    __shared__ unsigned long long hist_fx[HL];            // 2^24 fixed point: order-independent sums
    __shared__ float dst[d * d * nb];
    __shared__ float red[4];
    const Kp k = kps[blockIdx.x];
    int octave = k.octave & 255, layer = (k.octave >> 8) & 255;
    octave = octave < 128 ? octave : (-128 | octave);
    const float scale = octave >= 0 ? 1.0f / (float)(1 << octave) : (float)(1 << -octave);
    const float size = k.size * scale;
    const float ptx = k.x * scale, pty = k.y * scale;
    const OctaveGauss& G = octs[octave + 1];
    const float* img = G.g[layer];
    float ori = 360.0f - k.angle;
    if (fabsf(ori - 360.0f) < 1.1920929e-07f) ori = 0.0f;
    const float scl = size * 0.5f;
    const int px = __float2int_rn(ptx), py = __float2int_rn(pty);
    float cos_t = cosf(ori * (float)(3.14159265358979323846 / 180.0)), sin_t = sinf(ori * (float)(3.14159265358979323846 / 180.0));
    const float bins_per_rad = (float)nb / 360.0f, exp_scale = -1.0f / ((float)(d * d) * 0.5f);
    const float hist_width = kDescSclFctr * scl;
    int radius = __float2int_rn(hist_width * 1.4142135623730951f * (float)(d + 1) * 0.5f);
    radius = min(radius, (int)sqrt((double)G.w * G.w + (double)G.h * G.h));
    cos_t /= hist_width; sin_t /= hist_width;
    for (int i = threadIdx.x; i < HL; i += blockDim.x) hist_fx[i] = 0ull;
    __syncthreads();
    const int side = 2 * radius + 1;
    for (int s = threadIdx.x; s < side * side; s += blockDim.x) {
        const int i = s / side - radius, j = s % side - radius;
        const float c_rot = (float)j * cos_t - (float)i * sin_t, r_rot = (float)j * sin_t + (float)i * cos_t;
        const float rbin = r_rot + (float)(d / 2) - 0.5f, cbin = c_rot + (float)(d / 2) - 0.5f;
        const int r = py + i, c = px + j;
        if (!(rbin > -1.0f && rbin < (float)d && cbin > -1.0f && cbin < (float)d && r > 0 && r < G.h - 1 && c > 0 && c < G.w - 1)) continue;
        const float dx = img[(size_t)r * G.w + c + 1] - img[(size_t)r * G.w + c - 1];
        const float dy = img[(size_t)(r - 1) * G.w + c] - img[(size_t)(r + 1) * G.w + c];
        const float wgt = expf((c_rot * c_rot + r_rot * r_rot) * exp_scale);
        const float ang = fast_atan2_deg(dy, dx), mag = sqrtf(dx * dx + dy * dy) * wgt;
        float obin = (ang - ori) * bins_per_rad;
        const int r0 = (int)floorf(rbin), c0 = (int)floorf(cbin);
        int o0 = (int)floorf(obin);
        const float fr = rbin - (float)r0, fc = cbin - (float)c0, fo = obin - (float)o0;
        if (o0 < 0) o0 += nb;
        if (o0 >= nb) o0 -= nb;
        const float v_r1 = mag * fr, v_r0 = mag - v_r1;
        const float v_rc11 = v_r1 * fc, v_rc10 = v_r1 - v_rc11, v_rc01 = v_r0 * fc, v_rc00 = v_r0 - v_rc01;
        const float v111 = v_rc11 * fo, v110 = v_rc11 - v111, v101 = v_rc10 * fo, v100 = v_rc10 - v101;
        const float v011 = v_rc01 * fo, v010 = v_rc01 - v011, v001 = v_rc00 * fo, v000 = v_rc00 - v001;
        const int idx = ((r0 + 1) * (d + 2) + c0 + 1) * (nb + 2) + o0;
        auto add = [&](int at, float v) { atomicAdd(&hist_fx[at], (unsigned long long)(long long)llrintf(v * 16777216.0f)); };
        add(idx, v000); add(idx + 1, v001); add(idx + (nb + 2), v010); add(idx + (nb + 3), v011);
        add(idx + (d + 2) * (nb + 2), v100); add(idx + (d + 2) * (nb + 2) + 1, v101);
        add(idx + (d + 3) * (nb + 2), v110); add(idx + (d + 3) * (nb + 2) + 1, v111);
    }
    __syncthreads();
    // finalise: circular orientation bins, then normalise / clip / renormalise / quantise
    const int t = threadIdx.x;                                     // 128 threads = 4 x 4 x 8 outputs
    {
        const int i = t / (d * nb), j = (t / nb) % d, kk = t % nb;
        const int idx = ((i + 1) * (d + 2) + (j + 1)) * (nb + 2);
        float v = (float)(long long)hist_fx[idx + kk] * (1.0f / 16777216.0f);
        if (kk == 0) v += (float)(long long)hist_fx[idx + nb] * (1.0f / 16777216.0f);
        if (kk == 1) v += (float)(long long)hist_fx[idx + nb + 1] * (1.0f / 16777216.0f);
        dst[t] = v;
    }
    __syncthreads();
    auto block_sum = [&](float v) {
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
        if ((t & 31) == 0) red[t >> 5] = v;
        __syncthreads();
        const float r = (red[0] + red[1]) + (red[2] + red[3]);
        __syncthreads();
        return r;
    };
    const float nrm2 = block_sum(dst[t] * dst[t]);
    const float thr = sqrtf(nrm2) * kDescMagThr;
    const float v = fminf(dst[t], thr);
    const float nrm2b = block_sum(v * v);
    const float f = kIntDescrFctr / fmaxf(sqrtf(nrm2b), 1.1920929e-07f);
    const int q = __float2int_rn(v * f);
    desc[(size_t)blockIdx.x * 128 + t] = (float)max(0, min(255, q));
}

// retainBest keeps the nfeatures strongest keypoints; a dense 12 Mpx frame yields over a million.  Instead of moving and sorting
// them all on the host, a histogram over the top 16 bits of the responses (positive floats: integer order == float order) gives
// a conservative cut, only the keypoints above it travel, and the exact filter runs on those.
__global__ void __launch_bounds__(256) ks_resp_hist(const Kp* __restrict__ kps, int n, int* __restrict__ hist)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicAdd(&hist[__float_as_uint(kps[i].response) >> 16], 1);
}
__global__ void __launch_bounds__(256) ks_resp_select(const Kp* __restrict__ kps, int n, unsigned int min_bin, Kp* __restrict__ out, int* __restrict__ n_out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && (__float_as_uint(kps[i].response) >> 16) >= min_bin) out[atomicAdd(n_out, 1)] = kps[i];
}

}  // namespace

struct uavm_sift {
    int w = 0, h = 0, nfeatures = 0, nl = 3, n_oct = 0;
    float contrast = 0.01f, edge = 20.0f, sigma = 1.6f;
    int ow[kMaxOctaves], oh[kMaxOctaves];
    float* d_pool = nullptr;
    float* gauss[kMaxOctaves][kMaxLayers + 3];
    float* dog[kMaxOctaves][kMaxLayers + 2];
    float* d_tmp = nullptr;                  // row-pass output
    uint8_t* d_img = nullptr;                // staged host image
    OctaveGauss* d_octs = nullptr;
    Cand* d_cand = nullptr; Kp* d_kp = nullptr; Kp* d_kp_final = nullptr; float* d_desc = nullptr; int* d_counts = nullptr;
    int cap = 1 << 22, desc_cap = 1 << 17;
    int* d_hist = nullptr;                   // 65536 bins over the top 16 bits of the (positive) responses
    float taps[kMaxLayers + 3][kMaxTaps]; int radius[kMaxLayers + 3];
};

static void gaussian_taps(double sigma, float* taps, int& radius)
{
    int ks = (int)lrint(sigma * 8 + 1) | 1;                           // GaussianBlur: cvRound(sigma * 4 * 2 + 1) | 1 for CV_32F
    radius = ks / 2;
    std::vector<double> k(ks);
    double sum = 0;
    for (int i = 0; i < ks; i++) { const double x = i - (ks - 1) * 0.5; k[i] = exp(-0.5 * x * x / (sigma * sigma)); sum += k[i]; }
    for (int t = 0; t <= radius; t++) taps[t] = (float)(k[radius + t] / sum);
}

extern "C" int uavm_sift_create(uavm_ctx* ctx, int img_w, int img_h, int nfeatures, int n_octave_layers, double contrast_threshold,
                                double edge_threshold, double sigma, uavm_sift** out)
{
    if (!ctx || !out || img_w < 8 || img_h < 8 || n_octave_layers < 1 || n_octave_layers > kMaxLayers - 3 + 3 || sigma <= 0) return UAVM_EINVAL;
    *out = nullptr;
    UAVM_CUDA(ctx, cudaSetDevice(ctx->device));
    uavm_sift* s = new uavm_sift();
    s->w = img_w; s->h = img_h; s->nfeatures = nfeatures; s->nl = n_octave_layers;
    s->contrast = (float)contrast_threshold; s->edge = (float)edge_threshold; s->sigma = (float)sigma;
    const int bw = 2 * img_w, bh = 2 * img_h;
    int n_oct = (int)lrint(log((double)std::min(bw, bh)) / log(2.0) - 2) + 1;       // ... - firstOctave, firstOctave = -1
    if (n_oct > kMaxOctaves) n_oct = kMaxOctaves;
    if (n_oct < 1) n_oct = 1;
    s->n_oct = n_oct;
    size_t total = 0;
    for (int o = 0; o < n_oct; o++) {
        s->ow[o] = o == 0 ? bw : s->ow[o - 1] / 2; s->oh[o] = o == 0 ? bh : s->oh[o - 1] / 2;
        if (s->ow[o] < 1 || s->oh[o] < 1) { s->n_oct = o; break; }
        total += (size_t)s->ow[o] * s->oh[o] * (2 * s->nl + 5);
    }
    n_oct = s->n_oct;
    total += (size_t)bw * bh * 2;                                                   // row-pass scratch + base before the first blur
    if (cudaMalloc(&s->d_pool, total * sizeof(float) + 1024) != cudaSuccess) { cudaGetLastError(); delete s; UAVM_SET_ERR(ctx, "sift: out of device memory"); return UAVM_EFAIL; }
    float* p = s->d_pool;
    for (int o = 0; o < n_oct; o++) {
        const size_t px = (size_t)s->ow[o] * s->oh[o];
        for (int i = 0; i < s->nl + 3; i++) { s->gauss[o][i] = p; p += px; }
        for (int i = 0; i < s->nl + 2; i++) { s->dog[o][i] = p; p += px; }
    }
    s->d_tmp = p;
    // incremental sigmas (buildGaussianPyramid) and the initial blur (createInitialImage)
    const double k = pow(2.0, 1.0 / s->nl);
    gaussian_taps(sqrt(std::max(sigma * sigma - 0.5 * 0.5 * 4, 0.01)), s->taps[0], s->radius[0]);
    for (int i = 1; i < s->nl + 3; i++) {
        const double sp = pow(k, (double)(i - 1)) * sigma, st = sp * k;
        gaussian_taps(sqrt(st * st - sp * sp), s->taps[i], s->radius[i]);
    }
    bool fail = cudaMalloc(&s->d_img, (size_t)img_w * img_h * 3) != cudaSuccess || cudaMalloc(&s->d_octs, sizeof(OctaveGauss) * kMaxOctaves) != cudaSuccess ||
                cudaMalloc(&s->d_cand, sizeof(Cand) * s->cap) != cudaSuccess || cudaMalloc(&s->d_kp, sizeof(Kp) * s->cap) != cudaSuccess ||
                cudaMalloc(&s->d_kp_final, sizeof(Kp) * s->cap) != cudaSuccess || cudaMalloc(&s->d_desc, sizeof(float) * 128 * (size_t)s->desc_cap) != cudaSuccess ||
                cudaMalloc(&s->d_counts, 4 * sizeof(int)) != cudaSuccess || cudaMalloc(&s->d_hist, 65536 * sizeof(int)) != cudaSuccess;
    if (fail) { cudaGetLastError(); cudaFree(s->d_pool); cudaFree(s->d_img); cudaFree(s->d_octs); cudaFree(s->d_cand); cudaFree(s->d_kp); cudaFree(s->d_kp_final); cudaFree(s->d_desc); cudaFree(s->d_counts); cudaFree(s->d_hist); delete s; return UAVM_EFAIL; }
    std::vector<OctaveGauss> og(kMaxOctaves);
    for (int o = 0; o < n_oct; o++) { og[o].w = s->ow[o]; og[o].h = s->oh[o]; for (int i = 0; i < s->nl + 3; i++) og[o].g[i] = s->gauss[o][i]; }
    UAVM_CUDA(ctx, cudaMemcpy(s->d_octs, og.data(), sizeof(OctaveGauss) * kMaxOctaves, cudaMemcpyHostToDevice));
    *out = s;
    return UAVM_OK;
}

extern "C" void uavm_sift_destroy(uavm_ctx* ctx, uavm_sift* s)
{
    if (!s) return;
    if (ctx) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); }
    cudaFree(s->d_pool); cudaFree(s->d_img); cudaFree(s->d_octs); cudaFree(s->d_cand); cudaFree(s->d_kp); cudaFree(s->d_kp_final); cudaFree(s->d_desc); cudaFree(s->d_counts); cudaFree(s->d_hist);
    delete s;
}

// KeyPointsFilter::removeDuplicatedSorted + retainBest, then the firstOctave = -1 rescale (SIFT_Impl::detectAndCompute)
static void filter_keypoints(std::vector<Kp>& v, int nfeatures)
{
    std::sort(v.begin(), v.end(), [](const Kp& a, const Kp& b) {
        if (a.x != b.x) return a.x < b.x;
        if (a.y != b.y) return a.y < b.y;
        if (a.size != b.size) return a.size > b.size;
        if (a.angle != b.angle) return a.angle < b.angle;
        if (a.response != b.response) return a.response > b.response;
        return a.octave > b.octave;
    });
    size_t j = 0;
    for (size_t i = 0; i < v.size(); i++) {
        if (j > 0 && v[j - 1].x == v[i].x && v[j - 1].y == v[i].y && v[j - 1].size == v[i].size && v[j - 1].angle == v[i].angle) continue;
        v[j++] = v[i];
    }
    v.resize(j);
    if (nfeatures > 0 && (int)v.size() > nfeatures) {
        std::vector<float> resp(v.size());
        for (size_t i = 0; i < v.size(); i++) resp[i] = v[i].response;
        std::nth_element(resp.begin(), resp.begin() + (nfeatures - 1), resp.end(), std::greater<float>());
        const float amb = resp[nfeatures - 1];
        // the reference keeps an (implementation-defined) set of nfeatures - 1 best plus every point with response >= the
        // ambiguous one: exactly the points with response >= amb
        size_t m = 0;
        for (size_t i = 0; i < v.size(); i++) if (v[i].response >= amb) v[m++] = v[i];
        v.resize(m);
    }
    for (Kp& k : v) {
        k.x *= 0.5f; k.y *= 0.5f; k.size *= 0.5f;
        k.octave = (k.octave & ~255) | ((k.octave - 1) & 255);
    }
}

// detect + compute for one image.  kp_out: cv::KeyPoint records (class_id = -1); desc_out: n x 128 floats (integer valued 0..255).
// Returns the number of keypoints in *n_out (UAVM_EINVAL with *n_out set when cap is too small).
extern "C" int uavm_sift_detect_and_compute(uavm_ctx* ctx, uavm_sift* s, const uint8_t* bgr, int step, int is_device,
                                            uavm_keypoint* kp_out, float* desc_out, int cap, int* n_out)
{
    if (!ctx || !s || !bgr || step < 3 * s->w || !n_out) return UAVM_EINVAL;
    UAVM_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const uint8_t* src = bgr; int sstep = step;
    if (!is_device) {
        UAVM_CUDA(ctx, cudaMemcpy2DAsync(s->d_img, (size_t)s->w * 3, bgr, (size_t)step, (size_t)s->w * 3, s->h, cudaMemcpyHostToDevice, st));
        src = s->d_img; sstep = s->w * 3;
    }
    UAVM_CUDA(ctx, cudaMemcpyToSymbolAsync(c_taps, s->taps, sizeof(s->taps), 0, cudaMemcpyHostToDevice, st));
    UAVM_CUDA(ctx, cudaMemcpyToSymbolAsync(c_radius, s->radius, sizeof(s->radius), 0, cudaMemcpyHostToDevice, st));
    UAVM_CUDA(ctx, cudaMemsetAsync(s->d_counts, 0, 4 * sizeof(int), st));
    const int bw = s->ow[0], bh = s->oh[0];
    float* base_pre = s->d_tmp + (size_t)bw * bh;
    auto grid2 = [](int w, int h) { return dim3((w + 255) / 256, h); };
    ks_init<<<grid2(bw, bh), 256, 0, st>>>(src, sstep, s->w, s->h, base_pre);
    UAVM_CHECK_LAUNCH(ctx);
    auto blur = [&](const float* in, float* outp, int w, int h, int ki) {
        ks_blur<true><<<grid2(w, h), 256, 0, st>>>(in, s->d_tmp, w, h, ki);
        ks_blur<false><<<grid2(w, h), 256, 0, st>>>(s->d_tmp, outp, w, h, ki);
        ctx->launches += 2;
    };
    blur(base_pre, s->gauss[0][0], bw, bh, 0);
    for (int o = 0; o < s->n_oct; o++) {
        const int w = s->ow[o], h = s->oh[o];
        if (o > 0) { ks_down2<<<grid2(w, h), 256, 0, st>>>(s->gauss[o - 1][s->nl], s->ow[o - 1], s->gauss[o][0], w, h); ctx->launches++; }
        for (int i = 1; i < s->nl + 3; i++) blur(s->gauss[o][i - 1], s->gauss[o][i], w, h, i);
        const size_t px = (size_t)w * h;
        for (int i = 0; i < s->nl + 2; i++) { ks_sub<<<(unsigned)((px + 255) / 256), 256, 0, st>>>(s->gauss[o][i + 1], s->gauss[o][i], s->dog[o][i], px); ctx->launches++; }
        if (w > 2 * kImgBorder && h > 2 * kImgBorder) {
            OctaveDog D; D.w = w; D.h = h;
            for (int i = 0; i < s->nl + 2; i++) D.dog[i] = s->dog[o][i];
            const int threshold = (int)floor(0.5 * s->contrast / s->nl * 255);
            dim3 g((w - 2 * kImgBorder + 255) / 256, h - 2 * kImgBorder, s->nl);
            ks_extrema<<<g, 256, 0, st>>>(D, s->nl, o, threshold, s->contrast, s->edge, s->sigma, s->d_cand, s->d_counts, s->cap);
            UAVM_CHECK_LAUNCH(ctx);
        }
    }
    if (cudaGetLastError() != cudaSuccess) { UAVM_SET_ERR(ctx, "sift: kernel launch failed"); return UAVM_EFAIL; }
    int counts[2] = {0, 0};
    UAVM_CUDA(ctx, cudaMemcpyAsync(counts, s->d_counts, sizeof(int), cudaMemcpyDeviceToHost, st));
    UAVM_CUDA(ctx, cudaStreamSynchronize(st));
    if (counts[0] > s->cap) { UAVM_SET_ERR(ctx, "sift: %d extrema exceed the candidate buffer (%d)", counts[0], s->cap); return UAVM_EFAIL; }
    int n_cand = counts[0];
    if (n_cand > 0) {
        ks_orient<<<(n_cand * 32 + 255) / 256, 256, 0, st>>>(s->d_cand, n_cand, s->d_octs, s->d_kp, s->d_counts + 1, s->cap);
        UAVM_CHECK_LAUNCH(ctx);
    }
    UAVM_CUDA(ctx, cudaMemcpyAsync(counts, s->d_counts, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
    UAVM_CUDA(ctx, cudaStreamSynchronize(st));
    if (counts[1] > s->cap) { UAVM_SET_ERR(ctx, "sift: %d keypoints exceed the keypoint buffer (%d)", counts[1], s->cap); return UAVM_EFAIL; }
    const int n_kp = counts[1];
    std::vector<Kp> v;
    const int want = s->nfeatures > 0 ? s->nfeatures + s->nfeatures / 4 + 256 : 0;       // slack for duplicates removed before retainBest
    if (s->nfeatures > 0 && n_kp > 2 * want) {
        UAVM_CUDA(ctx, cudaMemsetAsync(s->d_hist, 0, 65536 * sizeof(int), st));
        ks_resp_hist<<<(n_kp + 255) / 256, 256, 0, st>>>(s->d_kp, n_kp, s->d_hist);
        UAVM_CHECK_LAUNCH(ctx);
        std::vector<int> hist(65536);
        UAVM_CUDA(ctx, cudaMemcpyAsync(hist.data(), s->d_hist, 65536 * sizeof(int), cudaMemcpyDeviceToHost, st));
        UAVM_CUDA(ctx, cudaStreamSynchronize(st));
        int bin = 65535, acc = 0;
        for (; bin > 0; bin--) { acc += hist[bin]; if (acc >= want) break; }
        ks_resp_select<<<(n_kp + 255) / 256, 256, 0, st>>>(s->d_kp, n_kp, (unsigned)bin, s->d_kp_final, s->d_counts + 2);
        UAVM_CHECK_LAUNCH(ctx);
        int n_sel = 0;
        UAVM_CUDA(ctx, cudaMemcpyAsync(&n_sel, s->d_counts + 2, sizeof(int), cudaMemcpyDeviceToHost, st));
        UAVM_CUDA(ctx, cudaStreamSynchronize(st));
        v.resize(n_sel);
        if (n_sel > 0) UAVM_CUDA(ctx, cudaMemcpy(v.data(), s->d_kp_final, sizeof(Kp) * n_sel, cudaMemcpyDeviceToHost));
        std::vector<Kp> probe(v);
        filter_keypoints(probe, s->nfeatures);
        if ((int)probe.size() >= s->nfeatures) v.swap(probe);
        else {                                                            // more duplicates than slack: exact filter on everything
            v.resize(n_kp);
            UAVM_CUDA(ctx, cudaMemcpy(v.data(), s->d_kp, sizeof(Kp) * n_kp, cudaMemcpyDeviceToHost));
            filter_keypoints(v, s->nfeatures);
        }
    } else {
        v.resize(n_kp);
        if (n_kp > 0) UAVM_CUDA(ctx, cudaMemcpy(v.data(), s->d_kp, sizeof(Kp) * n_kp, cudaMemcpyDeviceToHost));
        filter_keypoints(v, s->nfeatures);
    }
    const int n = (int)v.size();
    *n_out = n;
    if (n > cap) return UAVM_EINVAL;
    if (n > s->desc_cap) { UAVM_SET_ERR(ctx, "sift: %d keypoints exceed the descriptor buffer (%d); set nfeatures", n, s->desc_cap); return UAVM_EFAIL; }
    if (n == 0) return UAVM_OK;
    UAVM_CUDA(ctx, cudaMemcpyAsync(s->d_kp_final, v.data(), sizeof(Kp) * n, cudaMemcpyHostToDevice, st));
    ks_describe<<<n, 128, 0, st>>>(s->d_kp_final, n, s->d_octs, s->nl, s->d_desc);
    UAVM_CHECK_LAUNCH(ctx);
    if (desc_out) UAVM_CUDA(ctx, cudaMemcpyAsync(desc_out, s->d_desc, sizeof(float) * 128 * (size_t)n, cudaMemcpyDeviceToHost, st));
    UAVM_CUDA(ctx, cudaStreamSynchronize(st));
    if (kp_out)
        for (int i = 0; i < n; i++) {
            kp_out[i].x = v[i].x; kp_out[i].y = v[i].y; kp_out[i].size = v[i].size; kp_out[i].angle = v[i].angle; kp_out[i].response = v[i].response;
            kp_out[i].octave = v[i].octave; kp_out[i].class_id = -1;
        }
    return UAVM_OK;
}
