/* uavm.h — C ABI of libuavmosaic.so, the B200-native (sm_100a) hot path of UAV image mosaicking.
 *
 * Drop-in boundary for YuhuaXu/ImageMosaicing ("M/" = code/MosaicingCode/mosaicing/ of the
 * reference).  Each entry point names the reference call site it replaces.  Plain C types only:
 * pointers, sizes, PODs; no C++/torch/OpenCV types.  Return codes mirror the reference's
 * (M/MosaicWithoutPos.h:631): 0 ok, -1 invalid argument, -2 operation failed (CUDA error text via
 * uavm_last_error()).  There is no CPU fallback: every compute entry point fails with -2 if no
 * sm_100 device is available.
 *
 * Threading: one uavm_ctx per host thread / CUDA stream; objects created from a ctx are used
 * with that ctx.  All "host" pointers may be pageable; pinned memory makes uploads asynchronous:
 * a pinned source buffer handed to uavm_featureset_upload_* / uavm_canvas_set_image must stay valid
 * and unchanged until the next uavm_ctx_sync (or any result getter, which synchronises).
 */
#ifndef UAVM_H
#define UAVM_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define UAVM_OK 0
#define UAVM_EINVAL (-1)
#define UAVM_EFAIL (-2)

#define UAVM_DESC_DIM 128          /* SIFT-128 */
#define UAVM_MAX_CANDIDATES 400    /* maxNum, M/MosaicWithoutPos.cpp:5146 (9*44 = 396 used) */

/* ---- POD mirrors of the reference's wire types -------------------------------------------- */
typedef struct { float x, y; int32_t id; } uavm_sfpoint;                 /* SfPoint, M/Point.h:27-47 (12 B) */
typedef struct { int32_t queryIdx, trainIdx, imgIdx; float distance; } uavm_dmatch; /* cv::DMatch (16 B) */
typedef struct { float m[9]; } uavm_projectmat;                          /* ProjectMat, M/Bitmap.h:42-45 */
typedef struct { uavm_projectmat h; int32_t fixed; } uavm_imagetransform;/* ImageTransform, M/MosaicWithoutPos.h:224-228 */
typedef struct {                                                         /* MatchPointPairs, M/MosaicWithoutPos.h:135-153 (40 B) */
    uavm_sfpoint ptA; int32_t ptA_i; int32_t ptA_Fixed;
    uavm_sfpoint ptB; int32_t ptB_i; int32_t ptB_Fixed;
} uavm_matchpointpairs;
typedef struct {                                                         /* cv::KeyPoint of OpenCV 2.4 (28 B): record of keypoint_%d.key */
    float x, y, size, angle, response; int32_t octave, class_id;
} uavm_keypoint;
typedef struct {                                                         /* fields of IplImage the path reads (CV/core/types_c.h:460-493) */
    int32_t width, height, nChannels, widthStep;
    uint8_t* imageData;
} uavm_image;
typedef struct {                                                         /* UavMatchParam, M/MosaicWithoutPos.h:57-79 (hot-path fields) */
    float ransacDist;        /* 2.5 */
    int32_t blending;        /* 2 = multi-band */
    int32_t loadMatchPairs;  /* 0 */
    int32_t sampleTimes;     /* 1000 (Ransac2D default, M/mosaicimage.h:1735) */
    int32_t pairWindow;      /* 182: j in (i, min(n, i+pairWindow)), M/MosaicWithoutPos.cpp:5083 */
    int32_t minInnerPoints;  /* 30, :5049 */
    int32_t gridX, gridY;    /* 3, 3, :5041-5042 */
    int32_t maxNum;          /* 400, :5146 */
    float matchFrac;         /* 0.3, :5147 */
    int32_t numBands;        /* 5, :2179 */
    float overlapT;          /* 0.7, M/MosaicImage.cpp:2228 */
    uint32_t seed;           /* base seed of the RANSAC sample stream (replaces srand(time(0))) */
} uavm_param;
void uavm_param_default(uavm_param* p);

typedef struct {
    int32_t ok;            /* Ransac2D return value */
    int32_t n_inliers;
    int32_t max_support;   /* maxSupport */
    int32_t best_tuple;    /* index of the winning 4-tuple in the order of valid tuples, -1 none */
    int32_t n_tuples;      /* valid tuples consumed (realSamTimes) */
    int32_t n_counted;     /* hypotheses that passed the 5 px gate */
    float H[9];            /* refined homography, H[8] = max residual (M/LeastSquare.h:519) */
} uavm_ransac_result;

/* ---- context --------------------------------------------------------------------------------- */
typedef struct uavm_ctx uavm_ctx;
int uavm_ctx_create(int device, uavm_ctx** out);
void uavm_ctx_destroy(uavm_ctx* ctx);
int uavm_ctx_set_stream(uavm_ctx* ctx, void* cuda_stream);   /* cudaStream_t of the caller (e.g. torch's) */
int uavm_ctx_sync(uavm_ctx* ctx);
/* run independent stages concurrently: after fork() launches go to an internal side stream (ordered after all
 * work queued so far); unfork() returns to the main stream without waiting; join() orders the main stream
 * after the side work.  E.g. fork, ransac, unfork, warp, join: RANSAC (latency bound) overlaps the warp. */
int uavm_ctx_fork(uavm_ctx* ctx);
int uavm_ctx_unfork(uavm_ctx* ctx);
int uavm_ctx_join(uavm_ctx* ctx);
const char* uavm_last_error(const uavm_ctx* ctx);
int64_t uavm_ctx_launch_count(const uavm_ctx* ctx);          /* kernels launched so far by this ctx */
int uavm_ctx_sm_count(const uavm_ctx* ctx);

/* ---- device-resident feature set (replaces keypoint_%d.key / discriptor_%d.xml round trips,
 *      M/MosaicWithoutPos.cpp:4682-4734, :5073-5103) -------------------------------------------- */
typedef struct uavm_featureset uavm_featureset;
int uavm_featureset_create(uavm_ctx* ctx, int n_images, const int32_t* n_keypoints, uavm_featureset** out);
void uavm_featureset_destroy(uavm_ctx* ctx, uavm_featureset* fs);
/* desc: n x 128 (row-major); f32 descriptors must be integer-valued 0..255 (OpenCV SIFT output);
 * kp_xy: n x 2 floats (KeyPoint.pt).  is_device != 0: pointers are device memory. */
int uavm_featureset_upload_f32(uavm_ctx* ctx, uavm_featureset* fs, int image, const float* desc, const float* kp_xy, int is_device);
int uavm_featureset_upload_u8(uavm_ctx* ctx, uavm_featureset* fs, int image, const uint8_t* desc, const float* kp_xy, int is_device);

/* ---- SIFT on the GPU (csrc/sift.cu): replaces SiftExtraction_Thread's SIFT(2000, 3, 0.01, 20) detect + compute
 *      (M/MosaicWithoutPos.cpp:4852-4872).  One object per image size; kp_out are cv::KeyPoint records, desc_out n x 128 floats
 *      (integer valued 0..255, OpenCV's descriptor format).  nfeatures = 0 keeps every keypoint. */
typedef struct uavm_sift uavm_sift;
int uavm_sift_create(uavm_ctx* ctx, int img_w, int img_h, int nfeatures, int n_octave_layers, double contrast_threshold,
                     double edge_threshold, double sigma, uavm_sift** out);
void uavm_sift_destroy(uavm_ctx* ctx, uavm_sift* s);
int uavm_sift_detect_and_compute(uavm_ctx* ctx, uavm_sift* s, const uint8_t* bgr, int step, int is_device,
                                 uavm_keypoint* kp_out, float* desc_out, int cap, int* n_out);

/* ---- batched pair pipeline: match -> select -> RANSAC ------------------------------------------ */
typedef struct uavm_pairbatch uavm_pairbatch;
/* pair_ij: n_pairs x 2 (query image i, train image j), host memory. */
int uavm_pairbatch_create(uavm_ctx* ctx, uavm_featureset* fs, int n_pairs, const int32_t* pair_ij, uavm_pairbatch** out);
void uavm_pairbatch_destroy(uavm_ctx* ctx, uavm_pairbatch* pb);
/* K2: exact L2 1-NN of every query descriptor in the train image (replaces FlannBasedMatcher().match,
 * M/MosaicWithoutPos.cpp:5108-5110).  tcgen05 u8 x u8 -> s32 distance-GEMM with fused arg-min. */
int uavm_pairbatch_match(uavm_ctx* ctx, uavm_pairbatch* pb);
/* K3: sort by (d2, queryIdx) + nMatch = Min(maxNum, frac*size) + grid quota (replaces :5111, :5146-5153). */
int uavm_pairbatch_select(uavm_ctx* ctx, uavm_pairbatch* pb, int width, int height, int grid_x, int grid_y, int max_num, double frac);
/* K4: Ransac2D (replaces :5169 -> M/mosaicimage.h:1729-2035).  seeds: n_pairs host values, or NULL
 * for seed[p] = base_seed + p.  Sample stream = MSVC rand() LCG seeded per pair. */
int uavm_pairbatch_ransac(uavm_ctx* ctx, uavm_pairbatch* pb, float ransac_dist, int sample_times, const uint32_t* seeds, uint32_t base_seed);
/* results -> host (each call synchronises the ctx stream) */
int uavm_pairbatch_get_matches(uavm_ctx* ctx, uavm_pairbatch* pb, int pair, uavm_dmatch* out, int cap, int* n_out);
int uavm_pairbatch_get_candidates(uavm_ctx* ctx, uavm_pairbatch* pb, int pair, uavm_sfpoint* pts1, uavm_sfpoint* pts2, int cap, int* n_out);
int uavm_pairbatch_get_ransac(uavm_ctx* ctx, uavm_pairbatch* pb, int pair, uint8_t* inlier_mask, int cap, uavm_ransac_result* res);
/* accept rule nInliers > minInnerPoints and MatchPointPairs assembly (:5201-5221): appends the inliers of
 * every accepted pair in pair order.  Returns the number written in *n_out. */
int uavm_pairbatch_collect(uavm_ctx* ctx, uavm_pairbatch* pb, int min_inner_points, uavm_matchpointpairs* out, int cap, int* n_out, int* n_accepted_pairs);

/* ---- single-pair host-buffer seams (the calls the retained C++ host makes instead of OpenCV/own loops) */
/* replaces matcher.match(descriptors1, descriptors2, matches), :5108-5110 */
int uavm_match(uavm_ctx* ctx, const float* desc1, int n1, const float* desc2, int n2, uavm_dmatch* matches);
/* replaces std::sort + SelectMatchPairs, :5111, :5146-5153.  kp*_xy: n x 2. */
int uavm_select(uavm_ctx* ctx, const uavm_dmatch* matches, int n_matches, const float* kp1_xy, int n1, const float* kp2_xy, int n2,
                int width, int height, int grid_x, int grid_y, int max_num, double frac,
                uavm_sfpoint* pts1, uavm_sfpoint* pts2, int cap, int* n_out);
/* replaces Ransac2D(vecMatch1, vecMatch2, vecInner1, vecInner2, homo, ransacDist), :5169 */
int uavm_ransac2d(uavm_ctx* ctx, const uavm_sfpoint* pts1, const uavm_sfpoint* pts2, int n, float ransac_dist, int sample_times,
                  uint32_t seed, uavm_sfpoint* inner1, uavm_sfpoint* inner2, int cap, uavm_ransac_result* res);

/* ---- K8 global affine alignment (host; replaces BundleAdjustmentSparse, :6971-7202) ----------- */
int uavm_align_affine(const uavm_matchpointpairs* pairs, int n_pairs, const uavm_imagetransform* init, int n_images,
                      int n_fixed, uavm_imagetransform* out);
/* f3, the constrained variants (dead code in the shipped reference, methodType = 0 at :4588): the affine least squares with the
 * soft similarity constraints nC (a - d) = 0, nC (b + c) = 0 per free image (BundleAdjustmentSparseConstraint, :6032-6300), and the
 * rotation-constrained Gauss-Newton refinement that starts from it (SparseAffineRotConstraint, :6302-6808; the reference runs
 * 10 iterations; `weight` scales the per-image constraint weight int(nC * weight)). */
int uavm_align_affine_constrained(const uavm_matchpointpairs* pairs, int n_pairs, const uavm_imagetransform* init, int n_images,
                                  int n_fixed, uavm_imagetransform* out);
int uavm_align_affine_rot(const uavm_matchpointpairs* pairs, int n_pairs, const uavm_imagetransform* init, int n_images,
                          int n_fixed, float weight, int iterations, uavm_imagetransform* out);
/* largest connected component of the pair graph (Select_Connected_Matched_Images, :2754-2796) */
int uavm_connected_images(const uavm_matchpointpairs* pairs, int n_pairs, int n_images, int32_t* label);
/* stages [B] + [C] of MosaicWithoutPose on a match list (:4501-4652): connectivity, unconnected matches dropped, reference image 0
 * fixed, BundleAdjustmentSparse, unconnected images flagged m[8] = 0.  pairs is compacted in place (*n_used_out kept, fixed
 * flags set = the content of matchPairs.txt); label_out optional.  -2: nothing accepted / reference image unconnected. */
int uavm_global_align(uavm_matchpointpairs* pairs, int n_pairs, int n_images, uavm_imagetransform* transforms_out,
                      int32_t* label_out, int* n_used_out);

/* ---- warp + seam masks + multi-band blend (replaces LaplacianPyramidBlending, M/MosaicImage.cpp:2205-2510) */
typedef struct uavm_canvas uavm_canvas;
typedef struct {
    int32_t keep; int32_t beg_x, beg_y, chip_w, chip_h;
    float sx, sy; float quad[8]; float inv[9];
} uavm_chip_layout;
typedef struct { int32_t canvas_w, canvas_h; float dgx, dgy; } uavm_canvas_layout;
/* canvas sizing + per-image chip boxes (:2233-2348); H: n x 9 (already scaled), keep: n (1 = use) or NULL */
int uavm_canvas_layout_compute(const float* H, const int32_t* keep, int n, int img_w, int img_h,
                               uavm_canvas_layout* canvas, uavm_chip_layout* chips);
/* overlap filter (ResampleByOverlap, :2070-2201): keep[n] out */
int uavm_resample_by_overlap(const float* H, int n, int img_w, int img_h, float overlap_t, int32_t* keep);

int uavm_canvas_create(uavm_ctx* ctx, int n_images, int img_w, int img_h, const float* H, const int32_t* keep, uavm_canvas** out);
void uavm_canvas_destroy(uavm_ctx* ctx, uavm_canvas* cv);
int uavm_canvas_get_layout(uavm_canvas* cv, uavm_canvas_layout* canvas, uavm_chip_layout* chips);
/* multi-GPU canvas sharding: this ctx produces only the canvas rectangle [x0, x1) x [y0, y1) (even edges, or the canvas
 * size).  The blend computes exactly the pyramid regions the rectangle depends on, so the assembled rectangles are
 * bit-identical to the unsharded mosaic (at most 5 bands).  Chips that cannot contribute are deactivated
 * (uavm_canvas_is_active == 0: no need to set their image).  uavm_canvas_set_band = full-width rectangle; `halo` is ignored. */
int uavm_canvas_set_rect(uavm_ctx* ctx, uavm_canvas* cv, int x0, int y0, int x1, int y1);
int uavm_canvas_set_band(uavm_ctx* ctx, uavm_canvas* cv, int y0, int y1, int halo);
int uavm_canvas_is_active(uavm_canvas* cv, int image);
/* bytes per source pixel of the canvas' frame pool in HBM: 3 = frames are kept in the caller's BGR layout (width % 16 == 0: K5 reads
 * BGR taps, uavm_canvas_set_image is a plain copy), 4 = BGRA pool filled by a conversion kernel */
int uavm_canvas_source_layout(const uavm_canvas* cv);
/* source frame n (BGR u8 interleaved, `step` bytes per row); is_device != 0: device pointer */
int uavm_canvas_set_image(uavm_ctx* ctx, uavm_canvas* cv, int image, const uint8_t* bgr, int step, int is_device);
/* ---- JPEG frames decoded on the device (csrc/decode.cu; nvJPEG loaded at run time) — replaces the callers' cvLoadImage
 *      (M/mosaicing.cpp:51-100, M/MosaicWithoutPos.cpp:10224-10308).  backend: 0 default, 1 hybrid, 2 GPU hybrid, 3 the GPU's
 *      hardware JPEG engines (baseline single-scan streams, through the batched entry point). */
typedef struct uavm_jpeg uavm_jpeg;
int uavm_jpeg_create(uavm_ctx* ctx, int backend, uavm_jpeg** out);
void uavm_jpeg_destroy(uavm_ctx* ctx, uavm_jpeg* j);
int uavm_jpeg_info(uavm_ctx* ctx, uavm_jpeg* j, const uint8_t* jpeg, int64_t n_bytes, int* width, int* height);
int uavm_jpeg_decode_bgr(uavm_ctx* ctx, uavm_jpeg* j, const uint8_t* jpeg, int64_t n_bytes, uint8_t* d_bgr, int step, int width, int height);
/* source frame `image` from JPEG bytes: decoded straight into the canvas' BGR pool slot */
int uavm_canvas_set_image_jpeg(uavm_ctx* ctx, uavm_canvas* cv, uavm_jpeg* j, int image, const uint8_t* jpeg, int64_t n_bytes);
/* source frames [first, first + count) from JPEG bytes, decoded by several host threads at once (one nvJPEG decoder lane and CUDA
 * stream per thread; JPEG's entropy stage is sequential per frame).  uavm_jpeg_set_threads: n >= 1 threads, 0 = default (the
 * host's hardware threads, at most 32), -1 = nvJPEG's own batched decoder on the calling thread. */
int uavm_jpeg_set_threads(uavm_jpeg* j, int n);
int uavm_canvas_set_images_jpeg(uavm_ctx* ctx, uavm_canvas* cv, uavm_jpeg* j, int first, int count, const uint8_t* const* jpegs, const int64_t* n_bytes);
/* hardware JPEG engines nvJPEG reports for the device (0: none) */
int uavm_jpeg_hw_engines(uavm_jpeg* j);
/* device address of a source frame in a BGR pool (e.g. to run uavm_sift_detect_and_compute on it with is_device = 1) */
int uavm_canvas_image_ptr(uavm_canvas* cv, int image, const uint8_t** d_bgr, int* step);
/* K5: bilinear warp of every kept frame into its chip + validity mask (:2350-2448) */
int uavm_canvas_warp(uavm_ctx* ctx, uavm_canvas* cv);
/* same for images [first, first + count) only: a caller that streams frames in (uavm_canvas_set_image copies host
 * frames on an internal copy stream) warps each group as soon as it is set, overlapping PCIe with compute */
int uavm_canvas_warp_range(uavm_ctx* ctx, uavm_canvas* cv, int first, int count);
/* K5 for a mosaic: after uavm_canvas_seam_masks (which does not need the chips), warp only the chip pixels a blend of up to 5
 * bands can read — about a third of the work under seam masks; the mosaic is bit-identical, uavm_canvas_get_chip then returns
 * partly filled chips; a plain uavm_canvas_warp restores whole chips */
int uavm_canvas_warp_for_blend(uavm_ctx* ctx, uavm_canvas* cv);
/* K6: FindMasksByDistMap (:1761-1881); may run before the warp */
int uavm_canvas_seam_masks(uavm_ctx* ctx, uavm_canvas* cv);
/* K7: MultiBandBlender prepare/feed/blend + convertTo(CV_8U) (:2296-2299, :2476-2486) */
int uavm_canvas_blend(uavm_ctx* ctx, uavm_canvas* cv, int num_bands);
/* blending != 2 variant (MosaicImagesRefined, M/MosaicWithoutPos.cpp:2194-2352): last image wins */
int uavm_canvas_paste(uavm_ctx* ctx, uavm_canvas* cv);
/* copies -> host */
int uavm_canvas_get_chip(uavm_ctx* ctx, uavm_canvas* cv, int image, uint8_t* chip_bgr, int chip_step, uint8_t* mask, int mask_step);
int uavm_canvas_result_size(uavm_canvas* cv, int* w, int* h);   /* size of the blended / pasted result */
int uavm_canvas_get_result(uavm_ctx* ctx, uavm_canvas* cv, uint8_t* bgr, int step, uint8_t* mask, int mask_step);
/* rows [y0, y1) of the result, dense (canvas_w * 3 bytes per row); is_device != 0: dst is device memory (stream-ordered
 * copy, e.g. the send buffer of the NCCL gather of canvas bands) */
int uavm_canvas_copy_result_rows(uavm_ctx* ctx, uavm_canvas* cv, int y0, int y1, uint8_t* dst, int is_device);
/* rectangle [x0, x1) x [y0, y1) of the result, dense ((x1 - x0) * 3 bytes per row) */
int uavm_canvas_copy_result_rect(uavm_ctx* ctx, uavm_canvas* cv, int x0, int y0, int x1, int y1, uint8_t* dst, int is_device);

/* ---- multi-GPU (csrc/dist.cu): one uavm_ctx per GPU / rank, NCCL over NVLink, loaded at run time (libnccl.so.2).
 *      Replaces the worker fan-out of GetMatchedPairsOneToAllSIFT_MultiThread + PushMatchPairs
 *      (M/MosaicWithoutPos.cpp:5244-5295, :10137-10145): image pairs are sharded round-robin (pair p on rank p % world, local
 *      index p / world), ranks compute independently, one collective merges the results.  A multi-threaded C++ host calls
 *      uavm_dist_init from each of its GPU threads with the same id; torchrun ranks exchange the id over their store. */
typedef struct uavm_dist uavm_dist;
#define UAVM_DIST_ID_BYTES 128
int uavm_dist_unique_id(uint8_t* id_out, int id_bytes);                       /* rank 0: ncclGetUniqueId */
int uavm_dist_init(uavm_ctx* ctx, int rank, int world, const uint8_t* id, int id_bytes, uavm_dist** out);
void uavm_dist_destroy(uavm_ctx* ctx, uavm_dist* d);
int uavm_dist_rank(const uavm_dist* d);
int uavm_dist_world(const uavm_dist* d);
/* all ranks: merged MatchPointPairs list of ALL n_pairs_global pairs (accepted pairs in global pair order) — identical on
 * every rank and identical to uavm_pairbatch_collect of one context holding every pair.  Device-side pack, ncclAllGather of
 * fixed-size per-pair records, device-side compaction, one device-to-host copy of the dense list.  out == NULL: counts only. */
int uavm_pairbatch_allgather(uavm_ctx* ctx, uavm_dist* d, uavm_pairbatch* pb, int n_pairs_global, int min_inner_points,
                             uavm_matchpointpairs* out, int cap, int* n_out, int* n_accepted_pairs);
/* rects: world x 4 (x0, y0, x1, y1), rank r blended rects[r] (uavm_canvas_set_rect); afterwards root's result is the whole
 * mosaic.  Each rank copies its rectangle straight into the root's mosaic buffer over NVLink (mapped with CUDA IPC; grouped
 * ncclSend / ncclRecv when the ranks share a process or IPC is unavailable).  Stream ordered; synchronise with uavm_ctx_sync. */
int uavm_canvas_gather(uavm_ctx* ctx, uavm_dist* d, uavm_canvas* cv, const int32_t* rects, int root);
/* Fused blend + gather (collective; once per canvas, on every rank, after uavm_canvas_set_rect and before uavm_canvas_blend): the
 * root's mosaic buffer is mapped into the other ranks (CUDA IPC) and the level-0 kernel of their blend stores its rectangle there
 * as well, over NVLink, tile by tile while it computes; uavm_canvas_gather is then only the completion barrier.  If the mapping is
 * not possible nothing is bound and uavm_canvas_gather copies as before.  One canvas per uavm_dist is bound at a time: binding
 * another one, or destroying the canvas or the uavm_dist, unbinds. */
int uavm_canvas_bind_root(uavm_ctx* ctx, uavm_dist* d, uavm_canvas* cv, int root);
int uavm_canvas_bound_root(const uavm_canvas* cv);      /* root of that binding, -1: none */
int uavm_dist_broadcast(uavm_ctx* ctx, uavm_dist* d, void* device_buf, int64_t bytes, int root);

/* ---- top-level shim with the shape of MosaicVavImages (M/MosaicWithoutPos.h:638-645,
 *      M/MosaicWithoutPos.cpp:10148-10214).  Features are supplied by the caller (SIFT extraction is
 *      upstream of this library): desc[i] n_kp[i] x 128 f32, kp_xy[i] n_kp[i] x 2.
 *      result is allocated by the library (uavm_free); returns 0 / -1 / -2 like the reference. */
int uavm_mosaic_images(uavm_ctx* ctx, const uavm_image* images, int n_images,
                       const float* const* desc, const float* const* kp_xy, const int32_t* n_kp,
                       const uavm_param* param, float scale,
                       uavm_image* result, int* num_mosaiced, uavm_imagetransform* transforms_out);
/* MosaicVavImages' own signature (M/MosaicWithoutPos.h:638-645): images in, mosaic out; features from the GPU SIFT with the
 * reference's parameters SIFT(2000, 3, 0.01, 20).  pairs_out / n_pairs_out optional (uavm_free). */
int uavm_mosaic_images_sift(uavm_ctx* ctx, const uavm_image* images, int n_images, const uavm_param* param, float scale,
                            uavm_image* result, int* num_mosaiced, uavm_imagetransform* transforms_out,
                            uavm_matchpointpairs** pairs_out, int* n_pairs_out);
/* the loadMatchPairs = 1 path of MosaicWithoutPose (M/MosaicWithoutPos.cpp:4465-4477), without the stdin prompt: the match
 * list (e.g. uavm_match_file_read of a matchPairs.match written by the original tool or by uavm_match_file_write) replaces
 * feature matching; connectivity -> global alignment -> warp / blend run as in uavm_mosaic_images. */
int uavm_mosaic_from_matches(uavm_ctx* ctx, const uavm_image* images, int n_images,
                             const uavm_matchpointpairs* pairs, int n_pairs, const uavm_param* param, float scale,
                             uavm_image* result, int* num_mosaiced, uavm_imagetransform* transforms_out);
/* same as uavm_mosaic_images, and also returns the accepted inlier matches (what the reference writes to
 * feature_temp/matchPairs.match, :4492): *pairs_out is allocated by the library (uavm_free) */
int uavm_mosaic_images_ex(uavm_ctx* ctx, const uavm_image* images, int n_images,
                          const float* const* desc, const float* const* kp_xy, const int32_t* n_kp,
                          const uavm_param* param, float scale,
                          uavm_image* result, int* num_mosaiced, uavm_imagetransform* transforms_out,
                          uavm_matchpointpairs** pairs_out, int* n_pairs_out);
/* the chunk loop of MosaicUavVideo (M/MosaicWithoutPos.cpp:10252-10300) over an already decoded frame sequence: frames
 * [n, n + max_once_mosaic_num) are mosaicked independently, the next chunk starts at n + numMosaiced.  *results_out (and every
 * results_out[k].imageData) and *first_frame_out are allocated by the library (uavm_free); a failed chunk has imageData = NULL. */
int uavm_mosaic_sequence(uavm_ctx* ctx, const uavm_image* images, int n_images,
                         const float* const* desc, const float* const* kp_xy, const int32_t* n_kp,
                         const uavm_param* param, float scale, int max_once_mosaic_num,
                         uavm_image** results_out, int32_t** first_frame_out, int* n_results_out);
void uavm_free(void* p);

/* ---- the reference's on-disk artefacts (host only, csrc/formats_host.cpp) -------------------------------------------------
 * matchPairs.match: int32 n + n x 40 B MatchPointPairs (WriteMatchPairs / LoadMatchPairs, M/MosaicWithoutPos.cpp:4736-4797) */
int uavm_match_file_count(const char* path, int* n_out);
int uavm_match_file_read(const char* path, uavm_matchpointpairs* out, int cap, int* n_out);
int uavm_match_file_write(const char* path, const uavm_matchpointpairs* pairs, int n);
/* matchPairs.txt: "imgA xA yA fixedA imgB xB yB fixedB" per line (WriteMatchPairs_ASC2, :4751-4772); ids are not stored */
int uavm_match_text_write(const char* path, const uavm_matchpointpairs* pairs, int n);
int uavm_match_text_read(const char* path, uavm_matchpointpairs* out, int cap, int* n_out);
/* tran0.txt: rows for images 1..N-1, "m0 .. m7 fixed" (OutTransform, :2798-2818); read returns image 0 = identity, fixed */
int uavm_transform_file_write(const char* path, const uavm_imagetransform* t, int n_images);
int uavm_transform_file_read(const char* path, uavm_imagetransform* out, int cap, int* n_images_out);
/* ImportTransform's format (:2820-2843): count, then 9 floats per image */
int uavm_transform_import(const char* path, uavm_imagetransform* out, int cap, int* n_images_out);
/* keypoint_%d.key: int32 n + n x 28 B cv::KeyPoint; discriptor_%d.xml: FileStorage node "descriptor", CV_32F (:4682-4734) */
int uavm_key_file_read(const char* path, uavm_keypoint* out, int cap, int* n_out);
int uavm_key_file_write(const char* path, const uavm_keypoint* kp, int n);
int uavm_descriptor_xml_size(const char* path, int* rows, int* cols);
int uavm_descriptor_xml_read(const char* path, float* out, int cap_floats, int* rows, int* cols);
int uavm_descriptor_xml_write(const char* path, const float* desc, int rows, int cols);

#ifdef __cplusplus
}
#endif
#endif
