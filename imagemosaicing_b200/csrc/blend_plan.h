// blend_plan.h — host-side region planner of the multi-band blend (K7) and of canvas sharding.
//
// MultiBandBlender::feed (OpenCV stitching/blenders.cpp; driven by M/MosaicImage.cpp:2476-2480) builds, per image, a
// Laplacian pyramid of the image's padded ROI and accumulates it, weighted by the Gaussian pyramid of its mask, into the
// canvas pyramid.  With the seam masks of FindMasksByDistMap an image owns about 1 / coverage of its chip, so most of that
// work multiplies by a weight that is exactly zero.  The planner computes, per chip and pyramid level, the rectangle U_i where
// the weight can be non-zero inside this context's canvas rectangle, and the rectangle C_i of pyramid values those
// contributions depend on; the kernels only produce C_i and only evaluate U_i.  Skipped terms are exact zeros (dst +
// short(lap * 0) == dst, wsum + 0 == wsum), so the result is bit-identical to feeding whole ROIs.
//
// Dependencies (level sizes below the top level are even: ROIs are multiples of 2^bands):
//   pyrDown: level i+1 pixel x reads level i pixels 2x-2 .. 2x+2 (reflect-101 stays inside the clipped range)
//   pyrUp:   level i pixel x reads level i+1 pixels (x>>1)-1 .. (x>>1)+1 (reflect at the near edge, replicate at the far edge)
#pragma once
#include <algorithm>
#include <vector>

namespace uavm_plan {

constexpr int kMaxBands = 8;

struct IRect { int x0, y0, x1, y1; };                    // half open
inline bool is_empty(const IRect& r) { return r.x1 <= r.x0 || r.y1 <= r.y0; }
inline IRect make_empty() { return IRect{0, 0, 0, 0}; }
inline IRect isect(const IRect& a, const IRect& b)
{
    IRect r{std::max(a.x0, b.x0), std::max(a.y0, b.y0), std::min(a.x1, b.x1), std::min(a.y1, b.y1)};
    return is_empty(r) ? make_empty() : r;
}
inline IRect hull(const IRect& a, const IRect& b)
{
    if (is_empty(a)) return b;
    if (is_empty(b)) return a;
    return IRect{std::min(a.x0, b.x0), std::min(a.y0, b.y0), std::max(a.x1, b.x1), std::max(a.y1, b.y1)};
}
inline IRect shifted(const IRect& a, int dx, int dy) { return is_empty(a) ? a : IRect{a.x0 + dx, a.y0 + dy, a.x1 + dx, a.y1 + dy}; }
// level i-1 rectangle -> the level i pixels its pyrUp reads
inline IRect up_support(const IRect& r, int w, int h)
{
    if (is_empty(r)) return r;
    return isect(IRect{(r.x0 >> 1) - 1, (r.y0 >> 1) - 1, ((r.x1 - 1) >> 1) + 2, ((r.y1 - 1) >> 1) + 2}, IRect{0, 0, w, h});
}
// level i+1 rectangle -> the level i pixels its pyrDown reads
inline IRect down_support(const IRect& r, int w, int h)
{
    if (is_empty(r)) return r;
    return isect(IRect{2 * r.x0 - 2, 2 * r.y0 - 2, 2 * (r.x1 - 1) + 3, 2 * (r.y1 - 1) + 3}, IRect{0, 0, w, h});
}
// level i rectangle of non-zero weights -> level i+1 pixels whose pyrDown can be non-zero
inline IRect down_image(const IRect& r, int w, int h)
{
    if (is_empty(r)) return r;
    return isect(IRect{(r.x0 - 2 + 1) >> 1, (r.y0 - 2 + 1) >> 1, ((r.x1 + 1) >> 1) + 1, ((r.y1 + 1) >> 1) + 1}, IRect{0, 0, w, h});
}
inline IRect even_aligned(const IRect& r)               // grow to even edges (quads, 16-byte pixel pairs)
{
    if (is_empty(r)) return r;
    return IRect{r.x0 & ~1, r.y0 & ~1, (r.x1 + 1) & ~1, (r.y1 + 1) & ~1};
}

// The canvas-level kernels work on 4 x 2 pixel blocks aligned to multiples of 4 columns of the canvas level; a block that is
// only partly inside a rectangle still reads the pyrUp neighbourhood of both of its quads.
inline IRect block_grown(const IRect& r, int gx)
{
    if (is_empty(r)) return r;
    return IRect{r.x0 - gx, r.y0, r.x1 + gx, r.y1};
}
inline IRect block_aligned4(const IRect& r)
{
    if (is_empty(r)) return r;
    return IRect{r.x0 & ~3, r.y0 & ~1, (r.x1 + 3) & ~3, (r.y1 + 1) & ~1};
}

inline int pad_to(int v, int nb) { return v + ((1 << nb) - v % (1 << nb)) % (1 << nb); }

struct FeedRoi { int tlx, tly, width, height, top, left; };
// MultiBandBlender::feed: gap = 3 * 2^bands, corners aligned to 2^bands, shifted back inside the (padded) canvas
inline FeedRoi feed_roi(int tl_x, int tl_y, int cw, int ch, int W, int H, int nb)
{
    const int gap = 3 * (1 << nb);
    int tlx = tl_x - gap > 0 ? tl_x - gap : 0, tly = tl_y - gap > 0 ? tl_y - gap : 0;
    int brx = tl_x + cw + gap < W ? tl_x + cw + gap : W, bry = tl_y + ch + gap < H ? tl_y + ch + gap : H;
    tlx = (tlx >> nb) << nb; tly = (tly >> nb) << nb;
    const int width = pad_to(brx - tlx, nb), height = pad_to(bry - tly, nb);
    brx = tlx + width; bry = tly + height;
    const int dy = bry - H > 0 ? bry - H : 0, dx = brx - W > 0 ? brx - W : 0;
    tlx -= dx; tly -= dy;
    FeedRoi r; r.tlx = tlx; r.tly = tly; r.width = width; r.height = height; r.top = tl_y - tly; r.left = tl_x - tlx;
    return r;
}

// canvas pyramid geometry of one context: padded size, level sizes, and the level rectangles S_i the context computes
struct CanvasPlan {
    int nb = 0, W = 0, H = 0;
    int lw[kMaxBands + 1], lh[kMaxBands + 1];
    IRect S[kMaxBands + 1];
};
// out: the output rectangle (level 0, even edges or the canvas edge)
inline CanvasPlan plan_canvas(int canvas_w, int canvas_h, int nb, IRect out)
{
    CanvasPlan P; P.nb = nb;
    P.W = pad_to(canvas_w, nb); P.H = pad_to(canvas_h, nb);
    P.lw[0] = P.W; P.lh[0] = P.H;
    for (int i = 1; i <= nb; i++) { P.lw[i] = (P.lw[i - 1] + 1) / 2; P.lh[i] = (P.lh[i - 1] + 1) / 2; }
    P.S[0] = nb > 0 ? even_aligned(out) : out;
    for (int i = 1; i <= nb; i++) {
        P.S[i] = up_support(block_aligned4(P.S[i - 1]), P.lw[i], P.lh[i]);
        if (i < nb) P.S[i] = even_aligned(P.S[i]);
    }
    return P;
}

struct ChipPlan {
    bool active = false;
    FeedRoi roi;
    int pw[kMaxBands + 1], ph[kMaxBands + 1];
    IRect U[kMaxBands + 1];     // ROI level-i coordinates: where this chip contributes inside S_i
    IRect C[kMaxBands + 1];     // levels >= 1: pyramid values kept in scratch; level 0: ROI pixels read from the chip and its mask
};
// chip box (beg_x, beg_y, cw, ch) in canvas coordinates; A0: bounding box of the non-zero mask pixels in CHIP coordinates
inline ChipPlan plan_chip(int beg_x, int beg_y, int cw, int ch, IRect A0, const CanvasPlan& P)
{
    ChipPlan c; const int nb = P.nb;
    c.roi = feed_roi(beg_x, beg_y, cw, ch, P.W, P.H, nb);
    c.pw[0] = c.roi.width; c.ph[0] = c.roi.height;
    for (int i = 1; i <= nb; i++) { c.pw[i] = (c.pw[i - 1] + 1) / 2; c.ph[i] = (c.ph[i - 1] + 1) / 2; }
    IRect A = shifted(isect(A0, IRect{0, 0, cw, ch}), c.roi.left, c.roi.top);
    for (int i = 0; i <= nb; i++) {
        if (i > 0) A = down_image(A, c.pw[i], c.ph[i]);
        c.U[i] = isect(A, shifted(P.S[i], -(c.roi.tlx >> i), -(c.roi.tly >> i)));
        c.C[i] = make_empty();
        if (!is_empty(c.U[i])) c.active = true;
    }
    if (!c.active) return c;
    for (int i = nb; i >= 1; i--) {
        IRect r = hull(c.U[i], up_support(block_grown(c.U[i - 1], 2), c.pw[i], c.ph[i]));
        if (i < nb) r = hull(r, down_support(c.C[i + 1], c.pw[i], c.ph[i]));
        if (i < nb) r = even_aligned(r);
        c.C[i] = r;
    }
    c.C[0] = c.U[0];
    if (nb >= 1) c.C[0] = hull(c.C[0], down_support(c.C[1], c.pw[0], c.ph[0]));
    return c;
}

// 1-D: ROI range [a, b) relative to a chip of n pixels, read through BORDER_REFLECT -> chip pixels touched (half open)
inline void reflect_range(int a, int b, int n, int& lo, int& hi)
{
    lo = n; hi = 0;
    if (b <= a) return;
    const int ia = std::max(a, 0), ib = std::min(b, n);
    if (ib > ia) { lo = ia; hi = ib; }
    if (a < 0) { lo = 0; hi = std::max(hi, std::min(n, -a)); }                   // -1 -> 0 ... a -> -a-1
    if (b > n) { hi = n; lo = std::min(lo, std::max(0, 2 * n - b)); }            // n -> n-1 ... b-1 -> 2n-b
}

}  // namespace uavm_plan
