/* oracle_blend.c — CPU restatement of OpenCV's detail::MultiBandBlender (prepare / feed / blend) as the
 * reference drives it (M/MosaicImage.cpp:2296-2299, :2476-2486: CV_16SC3 chips, 8-bit masks, 5 bands,
 * weight_type CV_32F, no GPU), followed by convertTo(CV_8U).  TEST INFRASTRUCTURE ONLY.
 *
 * The blender is third-party code that is NOT under /root/reference (the reference links prebuilt OpenCV
 * 2.4.0 binaries, Readme.md:7, M/usingCV24.h:59-84).  This file restates the published algorithm of
 * modules/stitching/src/blenders.cpp and modules/imgproc/src/pyramids.cpp (2.4 series; unchanged in 4.x):
 *   prepare : canvas padded to a multiple of 2^bands; dst Laplacian pyramid (int16 x3) and weight pyramid (f32)
 *   feed    : ROI grown by gap = 3*2^bands, aligned to 2^bands; copyMakeBorder(REFLECT) of the image,
 *             copyMakeBorder(CONSTANT 0) of mask/255; Laplacian pyramid by pyrDown/pyrUp (int16, saturating
 *             subtract); Gaussian weight pyramid by pyrDown (f32); dst += short(src * w), wsum += w
 *   blend   : dst = short(dst / (wsum + 1e-5)); collapse by pyrUp + saturating add; crop; zero where
 *             wsum <= 1e-5
 * int16 kernels: pyrDown = separable [1 4 6 4 1], BORDER_REFLECT_101, (sum + 128) >> 8;
 *                pyrUp (dst = 2 x src) = even: s[i-1] + 6 s[i] + s[i+1], odd: 4 (s[i] + s[i+1]), reflect-101 at the
 *                near edge, replicate at the far edge, (sum + 32) >> 6.
 * Parity: pinned BIT FOR BIT against cv2 4.13 detail_MultiBandBlender, including the operation order of the f32 weight
 * pyrDown (tests/test_cpu_blend.py); parity with the exact 2.4.0 binary is UNPINNED (no mosaic image ships with the reference).
 */
#include "oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

static int reflect101(int p, int n)
{
    if (n == 1) return 0;
    while (p < 0 || p >= n) { if (p < 0) p = -p; else p = 2 * n - 2 - p; }
    return p;
}
static int reflect(int p, int n)      /* BORDER_REFLECT: fedcba|abcdefgh|hgfedcb */
{
    if (n == 1) return 0;
    while (p < 0 || p >= n) { if (p < 0) p = -p - 1; else p = 2 * n - 1 - p; }
    return p;
}
static short sat16(int v) { return (short)(v < -32768 ? -32768 : (v > 32767 ? 32767 : v)); }

void orc_pyr_down_s16(const int16_t* src, int w, int h, int ch, int16_t* dst)
{
    int dw = (w + 1) / 2, dh = (h + 1) / 2;
    int* row = (int*)malloc(sizeof(int) * (size_t)dw * ch * 5);
    for (int y = 0; y < dh; y++) {
        for (int k = 0; k < 5; k++) {
            int sy = reflect101(2 * y - 2 + k, h);
            const int16_t* s = src + (size_t)sy * w * ch;
            int* r = row + (size_t)k * dw * ch;
            for (int x = 0; x < dw; x++)
                for (int c = 0; c < ch; c++) {
                    int x0 = reflect101(2 * x - 2, w), x1 = reflect101(2 * x - 1, w), x2 = reflect101(2 * x, w);
                    int x3 = reflect101(2 * x + 1, w), x4 = reflect101(2 * x + 2, w);
                    r[x * ch + c] = s[x2 * ch + c] * 6 + (s[x1 * ch + c] + s[x3 * ch + c]) * 4 + s[x0 * ch + c] + s[x4 * ch + c];
                }
        }
        for (int i = 0; i < dw * ch; i++) {
            int v = row[2 * dw * ch + i] * 6 + (row[dw * ch + i] + row[3 * dw * ch + i]) * 4 + row[i] + row[4 * dw * ch + i];
            dst[(size_t)y * dw * ch + i] = sat16((v + 128) >> 8);
        }
    }
    free(row);
}

/* dst is exactly 2w x 2h (the blender always works on sizes divisible by 2^bands) */
void orc_pyr_up_s16(const int16_t* src, int w, int h, int ch, int16_t* dst, int dw, int dh)
{
    (void)dh;
    int* rows = (int*)malloc(sizeof(int) * (size_t)dw * ch * (size_t)h);
    for (int y = 0; y < h; y++) {
        const int16_t* s = src + (size_t)y * w * ch;
        int* r = rows + (size_t)y * dw * ch;
        for (int x = 0; x < w; x++)
            for (int c = 0; c < ch; c++) {
                int xm = (x == 0) ? (w > 1 ? 1 : 0) : x - 1;          /* reflect-101 at the near edge */
                int xp = (x == w - 1) ? w - 1 : x + 1;                  /* replicate at the far edge */
                r[(2 * x) * ch + c] = s[xm * ch + c] + s[x * ch + c] * 6 + s[xp * ch + c];
                r[(2 * x + 1) * ch + c] = (s[x * ch + c] + s[xp * ch + c]) * 4;
            }
    }
    for (int y = 0; y < h; y++) {
        int ym = (y == 0) ? (h > 1 ? 1 : 0) : y - 1;
        int yp = (y == h - 1) ? h - 1 : y + 1;
        const int* r0 = rows + (size_t)ym * dw * ch; const int* r1 = rows + (size_t)y * dw * ch; const int* r2 = rows + (size_t)yp * dw * ch;
        int16_t* d0 = dst + (size_t)(2 * y) * dw * ch; int16_t* d1 = d0 + (size_t)dw * ch;
        for (int i = 0; i < dw * ch; i++) {
            d0[i] = sat16((r0[i] + r1[i] * 6 + r2[i] + 32) >> 6);
            d1[i] = sat16(((r1[i] + r2[i]) * 4 + 32) >> 6);
        }
    }
    free(rows);
}

/* f32 pyrDown of the weight maps, in the EXACT operation order of OpenCV 4.13's imgproc/pyramids.cpp as shipped in the
 * opencv-python wheels (pyramids.cpp is not a CPU-dispatched file, so it runs the SSE baseline: 4 float lanes, v_muladd = mul
 * then add, no FMA).  Float addition is not associative, so which form produced an output column is part of the result:
 *   row pass    column x = 0 (left border) and columns past the vector loop:  ((s2*6 + (s1+s3)*4) + s0) + s4      (scalar loop)
 *               columns 1 <= x < 1 + 4*floor((width0-1)/4):                   s2*6 + ((s1+s3)*4 + (s0+s4))        (PyrDownVecH)
 *               with width0 = min((w-3)/2 + 1, dw)  (the columns that need no border lookup)
 *   column pass columns x < 4*floor(dw/4):   (((r1+r3)+r2)*4 + ((r0+r4)+(r2+r2))) * (1/256)                       (PyrDownVecV)
 *               the tail:                    (((r2*6 + (r1+r3)*4) + r0) + r4) * (1/256)                           (scalar loop)
 * Pinned bit for bit against cv2.pyrDown on random f32 images of 1..60 columns (tests/test_cpu_blend.py).  OpenCV 2.4.0 (the
 * reference's binary, unpinned) used the scalar row form everywhere and the same column forms with 8-column groups. */
void orc_pyr_down_f32(const float* src, int w, int h, float* dst)
{
    int dw = (w + 1) / 2, dh = (h + 1) / 2;
    int width0 = (w - 3) / 2 + 1; if (w < 3) width0 = 0; if (width0 > dw) width0 = dw;
    const int hvec_end = 1 + 4 * ((width0 - 1 > 0 ? width0 - 1 : 0) / 4);      /* columns [1, hvec_end) take the vector row form */
    const int vvec_end = 4 * (dw / 4);
    float* row = (float*)malloc(sizeof(float) * (size_t)dw * 5);
    for (int y = 0; y < dh; y++) {
        for (int k = 0; k < 5; k++) {
            int sy = reflect101(2 * y - 2 + k, h);
            const float* s = src + (size_t)sy * w;
            float* r = row + (size_t)k * dw;
            for (int x = 0; x < dw; x++) {
                int x0 = reflect101(2 * x - 2, w), x1 = reflect101(2 * x - 1, w), x2 = reflect101(2 * x, w);
                int x3 = reflect101(2 * x + 1, w), x4 = reflect101(2 * x + 2, w);
                if (x >= 1 && x < hvec_end) {
                    float t = (s[x1] + s[x3]) * 4.f; float u = s[x0] + s[x4]; t = t + u;
                    r[x] = s[x2] * 6.f + t;
                } else
                    r[x] = s[x2] * 6.f + (s[x1] + s[x3]) * 4.f + s[x0] + s[x4];
            }
        }
        for (int x = 0; x < dw; x++) {
            const float r0 = row[x], r1 = row[dw + x], r2 = row[2 * dw + x], r3 = row[3 * dw + x], r4 = row[4 * dw + x];
            float v;
            if (x < vvec_end) { float a = (r1 + r3) + r2; float b = (r0 + r4) + (r2 + r2); v = a * 4.f + b; }
            else v = r2 * 6.f + (r1 + r3) * 4.f + r0 + r4;
            dst[(size_t)y * dw + x] = v * (1.f / 256.f);
        }
    }
    free(row);
}

int orc_multiband_blend(const int16_t** chips16, const uint8_t** masks, const int32_t* tl_x, const int32_t* tl_y,
                        const int32_t* chip_w, const int32_t* chip_h, int n, int canvas_w, int canvas_h, int num_bands,
                        uint8_t* out, uint8_t* out_mask)
{
    /* prepare (blenders.cpp: MultiBandBlender::prepare) */
    double max_len = (double)(canvas_w > canvas_h ? canvas_w : canvas_h);
    int nb = (int)ceil(log(max_len) / log(2.0));
    if (num_bands < nb) nb = num_bands;
    int W = canvas_w, H = canvas_h;
    W += ((1 << nb) - W % (1 << nb)) % (1 << nb);
    H += ((1 << nb) - H % (1 << nb)) % (1 << nb);
    int lw[16], lh[16];
    int16_t* dlap[16]; float* dw[16];
    lw[0] = W; lh[0] = H;
    for (int i = 0; i <= nb; i++) {
        if (i > 0) { lw[i] = (lw[i - 1] + 1) / 2; lh[i] = (lh[i - 1] + 1) / 2; }
        dlap[i] = (int16_t*)calloc((size_t)lw[i] * lh[i] * 3, sizeof(int16_t));
        dw[i] = (float*)calloc((size_t)lw[i] * lh[i], sizeof(float));
    }
    /* feed */
    for (int k = 0; k < n; k++) {
        int gap = 3 * (1 << nb);
        int tlx = tl_x[k] - gap > 0 ? tl_x[k] - gap : 0, tly = tl_y[k] - gap > 0 ? tl_y[k] - gap : 0;
        int brx = tl_x[k] + chip_w[k] + gap < W ? tl_x[k] + chip_w[k] + gap : W;
        int bry = tl_y[k] + chip_h[k] + gap < H ? tl_y[k] + chip_h[k] + gap : H;
        tlx = (tlx >> nb) << nb; tly = (tly >> nb) << nb;
        int width = brx - tlx, height = bry - tly;
        width += ((1 << nb) - width % (1 << nb)) % (1 << nb);
        height += ((1 << nb) - height % (1 << nb)) % (1 << nb);
        brx = tlx + width; bry = tly + height;
        int dy = bry - H > 0 ? bry - H : 0, dx = brx - W > 0 ? brx - W : 0;
        tlx -= dx; brx -= dx; tly -= dy; bry -= dy;
        int top = tl_y[k] - tly, left = tl_x[k] - tlx;
        int cw = chip_w[k], chh = chip_h[k];
        int pw[16], ph[16];
        int16_t* pyr[16]; float* wp[16];
        pw[0] = width; ph[0] = height;
        pyr[0] = (int16_t*)malloc(sizeof(int16_t) * (size_t)width * height * 3);
        wp[0] = (float*)calloc((size_t)width * height, sizeof(float));
        for (int y = 0; y < height; y++) {
            int sy = reflect(y - top, chh);
            for (int x = 0; x < width; x++) {
                int sx = reflect(x - left, cw);
                for (int c = 0; c < 3; c++) pyr[0][((size_t)y * width + x) * 3 + c] = chips16[k][((size_t)sy * cw + sx) * 3 + c];
                int iy = y - top, ix = x - left;
                if (iy >= 0 && iy < chh && ix >= 0 && ix < cw)
                    wp[0][(size_t)y * width + x] = (float)masks[k][(size_t)iy * cw + ix] * (float)(1. / 255.);
            }
        }
        for (int i = 0; i < nb; i++) {
            pw[i + 1] = (pw[i] + 1) / 2; ph[i + 1] = (ph[i] + 1) / 2;
            pyr[i + 1] = (int16_t*)malloc(sizeof(int16_t) * (size_t)pw[i + 1] * ph[i + 1] * 3);
            wp[i + 1] = (float*)malloc(sizeof(float) * (size_t)pw[i + 1] * ph[i + 1]);
            orc_pyr_down_s16(pyr[i], pw[i], ph[i], 3, pyr[i + 1]);
            orc_pyr_down_f32(wp[i], pw[i], ph[i], wp[i + 1]);
        }
        for (int i = 0; i < nb; i++) {                     /* createLaplacePyr: pyr[i] -= pyrUp(pyr[i+1]) */
            int16_t* tmp = (int16_t*)malloc(sizeof(int16_t) * (size_t)pw[i] * ph[i] * 3);
            orc_pyr_up_s16(pyr[i + 1], pw[i + 1], ph[i + 1], 3, tmp, pw[i], ph[i]);
            for (size_t e = 0; e < (size_t)pw[i] * ph[i] * 3; e++) pyr[i][e] = sat16((int)pyr[i][e] - (int)tmp[e]);
            free(tmp);
        }
        int x_tl = tlx, y_tl = tly, x_br = brx, y_br = bry;
        for (int i = 0; i <= nb; i++) {
            for (int y = y_tl; y < y_br; y++)
                for (int x = x_tl; x < x_br; x++) {
                    int y_ = y - y_tl, x_ = x - x_tl;
                    float wv = wp[i][(size_t)y_ * pw[i] + x_];
                    for (int c = 0; c < 3; c++) {
                        int16_t* d = &dlap[i][((size_t)y * lw[i] + x) * 3 + c];
                        *d = (int16_t)(*d + (short)(pyr[i][((size_t)y_ * pw[i] + x_) * 3 + c] * wv));
                    }
                    dw[i][(size_t)y * lw[i] + x] += wv;
                }
            x_tl /= 2; y_tl /= 2; x_br /= 2; y_br /= 2;
        }
        for (int i = 0; i <= nb; i++) { free(pyr[i]); free(wp[i]); }
    }
    /* blend */
    for (int i = 0; i <= nb; i++)
        for (size_t e = 0; e < (size_t)lw[i] * lh[i]; e++) {
            float wv = dw[i][e] + 1e-5f;
            for (int c = 0; c < 3; c++) dlap[i][e * 3 + c] = (short)(dlap[i][e * 3 + c] / wv);
        }
    for (int i = nb; i > 0; i--) {
        int16_t* tmp = (int16_t*)malloc(sizeof(int16_t) * (size_t)lw[i - 1] * lh[i - 1] * 3);
        orc_pyr_up_s16(dlap[i], lw[i], lh[i], 3, tmp, lw[i - 1], lh[i - 1]);
        for (size_t e = 0; e < (size_t)lw[i - 1] * lh[i - 1] * 3; e++) dlap[i - 1][e] = sat16((int)tmp[e] + (int)dlap[i - 1][e]);
        free(tmp);
    }
    for (int y = 0; y < canvas_h; y++)
        for (int x = 0; x < canvas_w; x++) {
            int m = dw[0][(size_t)y * W + x] > 1e-5f;
            if (out_mask) out_mask[(size_t)y * canvas_w + x] = m ? 255 : 0;
            for (int c = 0; c < 3; c++) {
                int v = m ? dlap[0][((size_t)y * W + x) * 3 + c] : 0;
                out[((size_t)y * canvas_w + x) * 3 + c] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));   /* convertTo(CV_8U) */
            }
        }
    for (int i = 0; i <= nb; i++) { free(dlap[i]); free(dw[i]); }
    return 0;
}
