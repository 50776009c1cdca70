/* consumer.c — a plain C99 translation unit that includes include/uavm.h and calls host-side entry points of libuavmosaic.so:
 * proves that the boundary is C-clean (no C++ types, no default arguments, no references) and linkable from C.
 *   gcc -std=c99 -Wall -Wextra -pedantic -I include tests/c_consumer/consumer.c -L imagemosaicing_b200 -l:libuavmosaic.so
 * Two images; image 1 = image 0 shifted by (+10, -4): BundleAdjustmentSparse's answer is the affine map x' = x + 10, y' = y - 4. */
#include <math.h>
#include <stdio.h>
#include <string.h>
#include "uavm.h"

int main(void)
{
    uavm_param p;
    uavm_matchpointpairs m[6];
    uavm_imagetransform init[2], out[2];
    int32_t label[2];
    const float pts[6][2] = {{10, 10}, {200, 30}, {50, 300}, {400, 400}, {321, 77}, {90, 222}};
    int i, rc, n_used = 0;

    uavm_param_default(&p);
    if (p.sampleTimes != 1000 || p.numBands != 5 || p.pairWindow != 182) { printf("defaults differ\n"); return 1; }
    if (sizeof(uavm_matchpointpairs) != 40 || sizeof(uavm_sfpoint) != 12 || sizeof(uavm_imagetransform) != 40 || sizeof(uavm_keypoint) != 28) {
        printf("wire sizes differ\n"); return 1;
    }
    memset(m, 0, sizeof(m)); memset(init, 0, sizeof(init));
    for (i = 0; i < 6; i++) {
        m[i].ptA.x = pts[i][0] + 10.0f; m[i].ptA.y = pts[i][1] - 4.0f; m[i].ptA.id = i; m[i].ptA_i = 0; m[i].ptA_Fixed = 1;
        m[i].ptB.x = pts[i][0]; m[i].ptB.y = pts[i][1]; m[i].ptB.id = i; m[i].ptB_i = 1;
    }
    for (i = 0; i < 2; i++) { init[i].h.m[0] = init[i].h.m[4] = init[i].h.m[8] = 1.0f; }
    init[0].fixed = 1;
    rc = uavm_connected_images(m, 6, 2, label);
    if (rc != UAVM_OK || label[0] != 1 || label[1] != 1) { printf("connectivity failed\n"); return 1; }
    rc = uavm_align_affine(m, 6, init, 2, 1, out);
    if (rc != UAVM_OK) { printf("align failed: %d\n", rc); return 1; }
    if (fabs(out[1].h.m[0] - 1.0) > 1e-4 || fabs(out[1].h.m[4] - 1.0) > 1e-4 || fabs(out[1].h.m[2] - 10.0) > 1e-2 || fabs(out[1].h.m[5] + 4.0) > 1e-2) {
        printf("unexpected transform %g %g %g %g\n", out[1].h.m[0], out[1].h.m[4], out[1].h.m[2], out[1].h.m[5]); return 1;
    }
    for (i = 0; i < 6; i++) m[i].ptA_Fixed = 0;
    rc = uavm_global_align(m, 6, 2, out, label, &n_used);
    if (rc != UAVM_OK || n_used != 6 || fabs(out[1].h.m[2] - 10.0) > 1e-2) { printf("global_align failed\n"); return 1; }
    printf("c consumer ok\n");
    return 0;
}
