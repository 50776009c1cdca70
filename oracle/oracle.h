/* oracle.h — CPU oracle for the match -> select -> RANSAC -> align -> warp/blend hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the reference's algorithm
 * (YuhuaXu/ImageMosaicing, code/MosaicingCode/mosaicing/ = "M/").  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * liboracle.so; the product library (imagemosaicing_b200/libuavmosaic.so) never does.
 *
 * Parity pinning:
 *   - orc_ransac2d and its sub-functions are pinned bit-for-bit against the reference's own
 *     Ransac2D/SolveHomographyMatrix/NonlinearLeastSquareProjection2/InverseMatrix compiled in
 *     place from /root/reference (oracle/_ref/libref_ransac.so, tests/test_cpu_oracle.py).
 *   - orc_align is pinned by the reference's golden fixtures matchPairs.txt -> tran0.txt.
 *   - orc_match_l2 is pinned against cv2.BFMatcher(NORM_L2) (the reference's FLANN matcher is
 *     approximate + randomised: parity for FLANN itself is UNPINNED, contract = exact 1-NN).
 *   - orc_warp_chip / orc_seam_masks / the overlap filter are pinned byte-for-byte against the
 *     reference's own loops (M/MosaicImage.cpp:2350-2448, :1761-1881, :2070-2201) compiled in place
 *     (oracle/_ref) and against golden vectors generated from them (the .npz files under tests/golden).
 *   - orc_multiband_blend restates OpenCV's detail::MultiBandBlender (third-party, not under
 *     /root/reference; version 2.4.0 per Readme.md:7); pinned against cv2 4.13's blender.
 *
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off -msse2 -mfpmath=sse: scalar IEEE, no FMA)
 */
#ifndef UAVM_ORACLE_H
#define UAVM_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---------------- sample stream (replaces srand(time(0)), M/mosaicimage.h:1777) ------------- */
/* MSVC rand(): s = s*214013 + 2531011; return (s >> 16) & 0x7fff */
uint32_t orc_lcg_next(uint32_t* state);
int orc_set_threads(int n);   /* OpenMP threads for the parallel loops; returns the value in effect */

/* ---------------- K2: exact brute-force L2 1-NN (replaces FLANN, M/MosaicWithoutPos.cpp:5108) */
/* A: na x 128 u8, B: nb x 128 u8.  train_idx[i] = argmin_j |A_i - B_j|^2 (lowest j on ties),
 * d2[i] = that squared distance (exact integer).  DMatch.distance == sqrtf((float)d2). */
/* same results, AVX2 integer dot products (oracle_fast.c): the matcher of bench.py's CPU arm */
void orc_match_l2_fast(const uint8_t* A, int na, const uint8_t* B, int nb, int dim, int32_t* train_idx, int32_t* d2);
void orc_match_l2(const uint8_t* A, int na, const uint8_t* B, int nb, int dim,
                  int32_t* train_idx, int32_t* d2);

/* ---------------- K3: sort + 3x3 grid quota (M/MosaicWithoutPos.cpp:5111,5146-5153,4977-5028) */
/* matches are (queryIdx = i, trainIdx = train_idx[i], d2[i]) for i < n; order key (d2, queryIdx).
 * kp1/kp2: keypoint xy (n1 x 2 / n2 x 2 floats).  Outputs: up to max_num candidates
 * (xy1, id1, xy2, id2) in selection order.  Returns the number selected. */
int orc_select(const int32_t* train_idx, const int32_t* d2, int n,
               const float* kp1_xy, const float* kp2_xy,
               int width, int height, int grid_x, int grid_y, int max_num, double frac,
               float* out_xy1, int32_t* out_id1, float* out_xy2, int32_t* out_id2);

/* ---------------- K4: RANSAC homography (M/mosaicimage.h:1729-2035) ------------------------- */
typedef struct {
    int32_t best_tuple;     /* index (0-based, in order of valid 4-tuples) of the winning tuple, -1 none */
    int32_t best_t;         /* counted-hypothesis index t of the winner (maxSupportIndex) */
    int32_t max_support;    /* maxSupport */
    int32_t n_tuples;       /* valid 4-tuples drawn (realSamTimes of the last evaluated tuple) */
    int32_t n_counted;      /* hypotheses that passed the 5 px gate (final t) */
    int32_t n_refined;      /* hypotheses that took the Gauss-Newton refine */
    int32_t n_inv_fail;     /* InverseMatrix returned 0 inside the refine (stale-matrix quirk) */
    int32_t early_exit;     /* 1 if the 0.99 early exit fired */
    uint64_t rand_calls;    /* rand() calls consumed */
} orc_ransac_stats;

int orc_inverse_matrix(const float* src, int order, float* dst, float eps);
int orc_solve_homography(const float* xy1, const float* xy2, int n, float h[9]);
int orc_nls_projection2(const float* xy1, const float* xy2, int n, float out[9], const float init[9],
                        float stop, int* n_inv_fail);
/* returns 1/0 like Ransac2D; inlier_mask[n] (0/1), H[9] (H[8] = max residual), n_inliers. */
int orc_ransac2d(const float* xy1, const float* xy2, int n, float ransac_dist, int sample_times,
                 uint32_t seed, uint8_t* inlier_mask, float H[9], int* n_inliers,
                 orc_ransac_stats* stats);
/* evaluate one 4-tuple exactly as the loop body does: returns 0 rejected (>5px), 1 kept as is,
 * 2 refined; h[9] = hypothesis; *support = inlier count over all n points. */
int orc_ransac_eval_tuple(const float* xy1, const float* xy2, int n, const int32_t idx[4],
                          float thr2, float h[9], int* support);

/* ---------------- K8: global affine alignment (M/MosaicWithoutPos.cpp:6971-7202) ----------- */
/* pairs: n x {imgA, xA, yA, fixedA, imgB, xB, yB, fixedB} as doubles (matchPairs.txt columns).
 * fixed_img[nImages] (1 = parameters not solved), T0: nImages x 9 floats (initial transforms,
 * used for fixed images).  out: nImages x 9 floats.  Returns 0 ok, <0 error. */
int orc_align_affine(const double* pairs, int n_pairs, const int32_t* fixed_img, const float* T0,
                     int n_images, float* out);

/* ---------------- K5: canvas layout + chip warp (M/MosaicImage.cpp:2233-2448) --------------- */
typedef struct {
    int32_t keep;                 /* 0 = skipped (abandoned or m[8]==0) */
    int32_t beg_x, beg_y;         /* integer chip corner on the canvas (vecCorners) */
    int32_t chip_w, chip_h;
    float sx, sy;                 /* sub-pixel shift (begBox - begBox32F) */
    float quad[8];                /* quad corners in chip coordinates (vecRectPoints) */
    float inv[9];                 /* InverseMatrix(H,3,.,1e-12f) */
} orc_chip_layout;
typedef struct {
    int32_t canvas_w, canvas_h;
    float dgx, dgy;
} orc_canvas_layout;
/* H: n x 9 floats (already scaled), keep_in[n] (vecAbandonInd, 1 = keep). */
int orc_canvas_layout_compute(const float* H, const int32_t* keep_in, int n, int img_w, int img_h,
                              orc_canvas_layout* canvas, orc_chip_layout* chips);
/* src: img_h x src_step bytes, 3 channels interleaved.  chip: chip_h x chip_step (3ch), mask:
 * chip_h x mask_step (255 inside / 0 outside).  Chip bytes where mask==0 are written as 0
 * (the reference leaves them uninitialised, M/MosaicImage.cpp:2341). */
void orc_warp_chip(const uint8_t* src, int img_w, int img_h, int src_step,
                   const orc_canvas_layout* canvas, const orc_chip_layout* chip,
                   uint8_t* chip_px, int chip_step, uint8_t* mask, int mask_step);

/* blending != 2 variant: MosaicImagesRefined (M/MosaicWithoutPos.cpp:2194-2352), last image wins.
 * H: n x 9; srcs[n]: img_h x src_step BGR frames.  First call with out == NULL returns the canvas size. */
int orc_paste(const float* H, int n, int img_w, int img_h, const uint8_t** srcs, int src_step,
              int* out_w, int* out_h, uint8_t* out, int out_step);

/* ---------------- K6: distance-map seam masks (M/MosaicImage.cpp:1761-1881) ----------------- */
/* masks[n]: chip_h x mask_step[n] u8, in/out. */
int orc_seam_masks(uint8_t** masks, const int32_t* mask_step, const orc_chip_layout* chips,
                   int n_valid, int canvas_w, int canvas_h);

/* ---------------- K7: multi-band blend (OpenCV detail::MultiBandBlender, 2.4-series) -------- */
/* chips16[n]: chip_h x chip_w x 3 int16 (dense), masks[n]: chip_h x chip_w u8 (dense).
 * out: canvas_h x canvas_w x 3 u8 (dense), out_mask canvas_h x canvas_w u8. */
int orc_multiband_blend(const int16_t** chips16, const uint8_t** masks, const int32_t* tl_x,
                        const int32_t* tl_y, const int32_t* chip_w, const int32_t* chip_h, int n,
                        int canvas_w, int canvas_h, int num_bands, uint8_t* out, uint8_t* out_mask);
void orc_pyr_down_s16(const int16_t* src, int w, int h, int ch, int16_t* dst);   /* dst (w+1)/2 x (h+1)/2 */
void orc_pyr_up_s16(const int16_t* src, int w, int h, int ch, int16_t* dst, int dw, int dh);
void orc_pyr_down_f32(const float* src, int w, int h, float* dst);

#ifdef __cplusplus
}
#endif
#endif
