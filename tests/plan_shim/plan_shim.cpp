// Test shim: exposes the host-side region planner of the blend (imagemosaicing_b200/csrc/blend_plan.h, header only) to ctypes.
#include "blend_plan.h"
using namespace uavm_plan;

// out[0..3] = W, H, nb, 0; then per level i = 0..nb: lw, lh, S.x0, S.y0, S.x1, S.y1
extern "C" void shim_plan_canvas(int cw, int ch, int nb, int ox0, int oy0, int ox1, int oy1, int* out)
{
    const CanvasPlan P = plan_canvas(cw, ch, nb, IRect{ox0, oy0, ox1, oy1});
    out[0] = P.W; out[1] = P.H; out[2] = P.nb; out[3] = 0;
    for (int i = 0; i <= nb; i++) {
        int* o = out + 4 + 6 * i;
        o[0] = P.lw[i]; o[1] = P.lh[i]; o[2] = P.S[i].x0; o[3] = P.S[i].y0; o[4] = P.S[i].x1; o[5] = P.S[i].y1;
    }
}

// out[0..6] = active, tlx, tly, width, height, top, left; then per level: pw, ph, U (4), C (4)
extern "C" void shim_plan_chip(int cw, int ch, int nb, int ox0, int oy0, int ox1, int oy1, int beg_x, int beg_y, int chip_w, int chip_h,
                               int ax0, int ay0, int ax1, int ay1, int* out)
{
    const CanvasPlan P = plan_canvas(cw, ch, nb, IRect{ox0, oy0, ox1, oy1});
    const ChipPlan c = plan_chip(beg_x, beg_y, chip_w, chip_h, IRect{ax0, ay0, ax1, ay1}, P);
    out[0] = c.active; out[1] = c.roi.tlx; out[2] = c.roi.tly; out[3] = c.roi.width; out[4] = c.roi.height; out[5] = c.roi.top; out[6] = c.roi.left;
    for (int i = 0; i <= nb; i++) {
        int* o = out + 7 + 10 * i;
        o[0] = c.pw[i]; o[1] = c.ph[i];
        o[2] = c.U[i].x0; o[3] = c.U[i].y0; o[4] = c.U[i].x1; o[5] = c.U[i].y1;
        o[6] = c.C[i].x0; o[7] = c.C[i].y0; o[8] = c.C[i].x1; o[9] = c.C[i].y1;
    }
}
