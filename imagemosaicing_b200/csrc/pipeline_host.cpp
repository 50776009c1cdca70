// pipeline_host.cpp — uavm_mosaic_images: the MosaicVavImages-shaped shim (M/MosaicWithoutPos.cpp:10148-10214
// -> CMosaicByPose::MosaicWithoutPose :4430-4679) composing the stage entry points of this library.
#include <string.h>
#include <vector>
#include "internal.h"

extern "C" int uavm_mosaic_images(uavm_ctx* ctx, const uavm_image* images, int n_images,
                                  const float* const* desc, const float* const* kp_xy, const int32_t* n_kp,
                                  const uavm_param* param, float scale,
                                  uavm_image* result, int* num_mosaiced, uavm_imagetransform* transforms_out)
{
    (void)images; (void)n_images; (void)desc; (void)kp_xy; (void)n_kp; (void)param; (void)scale; (void)result;
    (void)num_mosaiced; (void)transforms_out;
    UAVM_SET_ERR(ctx, "uavm_mosaic_images: not implemented yet");
    return UAVM_EFAIL;
}
