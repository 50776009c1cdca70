// overlap_host.cpp — ResampleByOverlap (M/MosaicImage.cpp:2070-2201): drop image n1 when its warped quad
// overlaps an earlier kept quad by more than overlapT of its own area.  Sequential O(N^2) float geometry
// on the host, restated with the reference's helpers and their quirks:
//   AreaOfQuadrangle :1884-1931, IsPointOnLineSegmentOfTwoPoints :1934-1948, GetAllIntersecPoints :1951-1996,
//   IsPointInQuadrangle :1999-2027, GetPointsInOverlapRegion :2030-2067,
//   AngleofPoint360 M/ImageMath.cpp:9-54, LineOf2Points1 :88-103, ABCToPolar :144-176,
//   IntersectionPointOf2PolarLines :399-412, AngleofPoint M/imageMath.h:26-87.
// Transcendentals (atan/acos/sin/cos) come from the host libm, so the last ulp can differ from the
// reference's MSVC runtime: parity for this stage is unpinned (no fixture exercises it).
#include <math.h>
#include <algorithm>
#include <vector>
#include "internal.h"

namespace {

const float pi = 3.1415926f;                                  // M/Bitmap.h:54
struct Pt { float x, y; };

float dist2pts(float x1, float y1, float x2, float y2) { return sqrtf((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2)); }

void area_of_quadrangle(const Pt c[4], float& area)
{
    const float v02x = c[0].x - c[2].x, v02y = c[0].y - c[2].y;
    const float v13x = c[1].x - c[3].x, v13y = c[1].y - c[3].y;
    const float L1 = sqrtf(v02x * v02x + v02y * v02y), L2 = sqrtf(v13x * v13x + v13y * v13y);
    const float L01 = dist2pts(c[0].x, c[0].y, c[1].x, c[1].y), L02 = dist2pts(c[0].x, c[0].y, c[2].x, c[2].y);
    const float L12 = dist2pts(c[1].x, c[1].y, c[2].x, c[2].y), L03 = dist2pts(c[0].x, c[0].y, c[3].x, c[3].y);
    const float L23 = dist2pts(c[2].x, c[2].y, c[3].x, c[3].y);
    const float P1 = (L01 + L02 + L12) * 0.5f;
    const float S1 = sqrtf(P1 * (P1 - L01) * (P1 - L02) * (P1 - L12));
    const float P2 = (L02 + L03 + L23) * 0.5f;
    const float S2 = sqrtf(P2 * (P2 - L02) * (P2 - L03) * (P2 - L23));
    if ((fabsf(S1) < 1e-4) && (fabsf(S2) < 1e-4)) { area = 0; return; }
    if (L1 * L2 > 0) {
        const float cos_t = (v02x * v13x + v02y * v13y) / (L1 * L2);
        float theta = acosf(cos_t);
        if (theta < 0) theta = theta + 3.1415926f;
        area = (float)(0.5 * L1 * L2 * sinf(theta));
    } else area = 0;
}

bool on_segment(Pt p, Pt a, Pt b)
{
    const float d12 = dist2pts(a.x, a.y, b.x, b.y), d01 = dist2pts(p.x, p.y, a.x, a.y), d02 = dist2pts(p.x, p.y, b.x, b.y);
    return fabsf(d01 + d02 - d12) < 1e-4;
}

void line_of_2_points(float& a, float& b, float& c, float x1, float y1, float x2, float y2)
{
    if (fabs(x1 - x2) < 0.000001) { a = 1.0f; b = 0; c = -x1; }
    else { a = (y1 - y2) / (x1 - x2); b = -1.0f; c = y1 - a * x1; }
}

float angle_of_point(float x, float y)
{
    float angle = 0;
    if (x >= 0) {
        if (y >= 0) { if (x != 0) angle = atanf(y / x); else angle = (y != 0) ? pi / 2 : 0; }
        else { if (x != 0) angle = atanf(y / x); else angle = -pi / 2; }
    } else {
        if (y >= 0) angle = pi + atanf(y / x);
        else angle = atanf(y / x) - pi;
    }
    return angle;
}

void abc_to_polar(float a, float b, float c, float& rho, float& theta)
{
    float a1 = 0, xc = 0, yc = 0;
    rho = fabsf(a * 0 + b * 0 + c) / sqrtf(a * a + b * b);
    if (b == 0) { yc = 0; xc = -c / a; }
    if (a == 0) { xc = 0; yc = -c / b; }
    if ((a != 0) && (b != 0)) { a1 = -1.0f / a; xc = -c / (a - a1); yc = a1 * xc; }
    theta = angle_of_point(xc, yc);
    if (theta < 0) theta = theta + 2 * pi;
}

void intersect_polar(float rho1, float th1, float rho2, float th2, Pt& p)
{
    const float det = sinf(th2) * cosf(th1) - sinf(th1) * cosf(th2);
    p.x = (rho1 * sinf(th2) - rho2 * sinf(th1)) / det;
    p.y = (-rho1 * cosf(th2) + rho2 * cosf(th1)) / det;
}

void all_intersections(const Pt c1[4], const Pt c2[4], std::vector<Pt>& out)
{
    for (int n = 0; n < 4; n++) { out.push_back(c1[n]); out.push_back(c2[n]); }
    for (int n1 = 0; n1 < 4; n1++) {
        const int m1 = (n1 + 1) & 3;
        float A1, B1, C1, rho1, th1;
        line_of_2_points(A1, B1, C1, c1[n1].x, c1[n1].y, c1[m1].x, c1[m1].y);
        abc_to_polar(A1, B1, C1, rho1, th1);
        for (int n2 = 0; n2 < 4; n2++) {
            const int m2 = (n2 + 1) & 3;
            float A2, B2, C2, rho2, th2;
            line_of_2_points(A2, B2, C2, c2[n2].x, c2[n2].y, c2[m2].x, c2[m2].y);
            abc_to_polar(A2, B2, C2, rho2, th2);
            Pt p;
            intersect_polar(rho1, th1, rho2, th2, p);
            if (on_segment(p, c1[n1], c1[m1]) && on_segment(p, c2[n2], c2[m2])) out.push_back(p);
        }
    }
}

bool in_quadrangle(Pt p, const Pt c[4])
{
    float area4 = 0;
    area_of_quadrangle(c, area4);
    float acc = 0;
    for (int n = 0; n < 4; n++) {
        Pt t[4] = {c[n], c[(n + 1) & 3], p, p};
        float a;
        area_of_quadrangle(t, a);
        acc += a;
    }
    return fabsf(area4 - acc) < 0.2;
}

void angle_of_point_360(float x, float y, float& ang)
{
    if (x >= 0) {
        if (y >= 0) { if (x != 0) ang = atanf(y / x); else ang = pi / 2; }
        else { if (x != 0) ang = 2 * pi + atanf(y / x); else ang = 3 * pi / 2; }
    } else ang = pi + atanf(y / x);
}

struct Ang { float dist; int seq; bool operator<(const Ang& r) const { return dist < r.dist; } };

void points_in_overlap(const Pt c1[4], const Pt c2[4], const std::vector<Pt>& cand, std::vector<Pt>& out)
{
    std::vector<Pt> in;
    float cx = 0, cy = 0;
    for (size_t i = 0; i < cand.size(); i++)
        if (in_quadrangle(cand[i], c1) && in_quadrangle(cand[i], c2)) { in.push_back(cand[i]); cx += cand[i].x; cy += cand[i].y; }
    cx /= in.size(); cy /= in.size();
    std::vector<Ang> ang;
    for (size_t i = 0; i < in.size(); i++) { Ang a; a.seq = (int)i; a.dist = 0; angle_of_point_360(in[i].x - cx, in[i].y - cy, a.dist); ang.push_back(a); }
    std::sort(ang.begin(), ang.end());
    for (size_t i = 0; i < in.size(); i++) out.push_back(in[ang[i].seq]);
}

void quad_of(const float* m, int w, int h, Pt q[4])
{
    const float cx[4] = {0.0f, (float)(w - 1), (float)(w - 1), 0.0f}, cy[4] = {0.0f, 0.0f, (float)(h - 1), (float)(h - 1)};
    for (int i = 0; i < 4; i++) {
        q[i].x = (cx[i] * m[0] + cy[i] * m[1] + m[2]) / (cx[i] * m[6] + cy[i] * m[7] + m[8]);
        q[i].y = (cx[i] * m[3] + cy[i] * m[4] + m[5]) / (cx[i] * m[6] + cy[i] * m[7] + m[8]);
    }
}

}  // namespace

extern "C" int uavm_resample_by_overlap(const float* H, int n, int img_w, int img_h, float overlap_t, int32_t* keep)
{
    if (!H || n <= 0 || !keep || img_w < 2 || img_h < 2) return UAVM_EINVAL;
    for (int i = 0; i < n; i++) keep[i] = 1;
    for (int n1 = 1; n1 < n; n1++) {
        const float* m1 = H + (size_t)n1 * 9;
        if (m1[8] == 0) continue;
        Pt q1[4];
        quad_of(m1, img_w, img_h, q1);
        float area1 = 0;
        area_of_quadrangle(q1, area1);
        bool satisfied = true;
        for (int n2 = 0; n2 < n1; n2++) {
            if (keep[n2] == 0) continue;
            const float* m2 = H + (size_t)n2 * 9;
            if (m2[8] == 0) continue;
            Pt q2[4];
            quad_of(m2, img_w, img_h, q2);
            std::vector<Pt> cand, ov;
            all_intersections(q1, q2, cand);
            points_in_overlap(q1, q2, cand, ov);
            if (ov.size() == 3) ov.push_back(ov[2]);
            float area2 = 0;
            if (ov.size() == 4) area_of_quadrangle(ov.data(), area2);
            if (area2 / area1 > overlap_t) { satisfied = false; break; }
        }
        if (!satisfied) keep[n1] = 0;
    }
    keep[n - 1] = 1;                                   // the last image is always kept (:2198)
    return UAVM_OK;
}
