// blend.cu — K7 multi-band blend (placeholder until the kernels land in this file).
#include "canvas.h"

void uavm_blend_free(uavm_canvas* cv) { (void)cv; }

extern "C" int uavm_canvas_blend(uavm_ctx* ctx, uavm_canvas* cv, int num_bands)
{
    (void)cv; (void)num_bands;
    UAVM_SET_ERR(ctx, "uavm_canvas_blend: not implemented yet");
    return UAVM_EFAIL;
}
extern "C" int uavm_canvas_paste(uavm_ctx* ctx, uavm_canvas* cv)
{
    (void)cv;
    UAVM_SET_ERR(ctx, "uavm_canvas_paste: not implemented yet");
    return UAVM_EFAIL;
}
extern "C" int uavm_canvas_get_result(uavm_ctx* ctx, uavm_canvas* cv, uint8_t* bgr, int step, uint8_t* mask, int mask_step)
{
    (void)cv; (void)bgr; (void)step; (void)mask; (void)mask_step;
    UAVM_SET_ERR(ctx, "uavm_canvas_get_result: not implemented yet");
    return UAVM_EFAIL;
}
