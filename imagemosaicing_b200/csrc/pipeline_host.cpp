// pipeline_host.cpp — uavm_mosaic_images: the MosaicVavImages-shaped shim (M/MosaicWithoutPos.cpp:10148-10214
// -> CMosaicByPose::MosaicWithoutPose :4430-4679) composing the stage entry points of this library.
// SIFT extraction (SiftExtraction_Thread, :4832-4887) is upstream of this library: the caller passes the
// descriptors and keypoints it extracted.
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "internal.h"

static int check_images(const uavm_image* images, int n_images)
{
    const int w = images[0].width, h = images[0].height;              // size taken from image 0 (:10185-10186)
    for (int i = 0; i < n_images; i++)
        if (!images[i].imageData || images[i].nChannels != 3 || images[i].width != w || images[i].height != h || images[i].widthStep < 3 * w) return UAVM_EINVAL;
    return UAVM_OK;
}

static int mosaic_tail(uavm_ctx* ctx, const uavm_image* images, int n_images, std::vector<uavm_matchpointpairs>& matches, const uavm_param& P,
                       float scale, uavm_image* result, int* num_mosaiced, uavm_imagetransform* transforms_out);

extern "C" int uavm_mosaic_images_ex(uavm_ctx* ctx, const uavm_image* images, int n_images,
                                     const float* const* desc, const float* const* kp_xy, const int32_t* n_kp,
                                     const uavm_param* param_in, float scale,
                                     uavm_image* result, int* num_mosaiced, uavm_imagetransform* transforms_out,
                                     uavm_matchpointpairs** pairs_out, int* n_pairs_out)
{
    // argument checks of MosaicVavImages (:10157-10165): -1 on invalid input
    if (!ctx || !images || n_images < 2 || !desc || !kp_xy || !n_kp || !result) return UAVM_EINVAL;
    memset(result, 0, sizeof(*result));
    if (pairs_out) *pairs_out = nullptr;
    if (n_pairs_out) *n_pairs_out = 0;
    uavm_param P;
    if (param_in) P = *param_in; else uavm_param_default(&P);
    if (check_images(images, n_images) != UAVM_OK) return UAVM_EINVAL;
    const int w = images[0].width, h = images[0].height;
    for (int i = 0; i < n_images; i++)
        if (n_kp[i] < 0 || (n_kp[i] > 0 && (!desc[i] || !kp_xy[i]))) return UAVM_EINVAL;
    if (num_mosaiced) *num_mosaiced = 0;

    // ---- [A] feature matching: GetMatchedPairsOneToAllSIFT_MultiThread (:5244) ----
    uavm_featureset* fs = nullptr; uavm_pairbatch* pb = nullptr;
    int rc = uavm_featureset_create(ctx, n_images, n_kp, &fs);
    for (int i = 0; rc == UAVM_OK && i < n_images; i++) rc = uavm_featureset_upload_f32(ctx, fs, i, desc[i], kp_xy[i], 0);
    std::vector<int32_t> pair_ij;
    for (int i = 0; i < n_images; i++)                                // j in (i, min(n, i + 182)) (:5083-5084)
        for (int j = i + 1; j < n_images && j < i + P.pairWindow; j++) { pair_ij.push_back(i); pair_ij.push_back(j); }
    const int n_pairs = (int)pair_ij.size() / 2;
    std::vector<uavm_matchpointpairs> matches;
    if (rc == UAVM_OK && n_pairs > 0) {
        rc = uavm_pairbatch_create(ctx, fs, n_pairs, pair_ij.data(), &pb);
        if (rc == UAVM_OK) rc = uavm_pairbatch_match(ctx, pb);
        if (rc == UAVM_OK) rc = uavm_pairbatch_select(ctx, pb, w, h, P.gridX, P.gridY, P.maxNum, (double)P.matchFrac);
        if (rc == UAVM_OK) rc = uavm_pairbatch_ransac(ctx, pb, P.ransacDist, P.sampleTimes, nullptr, P.seed);
        int n_m = 0, n_acc = 0;
        if (rc == UAVM_OK) rc = uavm_pairbatch_collect(ctx, pb, P.minInnerPoints, nullptr, 0, &n_m, &n_acc);
        if (rc == UAVM_OK && n_m > 0) {
            matches.resize(n_m);
            rc = uavm_pairbatch_collect(ctx, pb, P.minInnerPoints, matches.data(), n_m, &n_m, &n_acc);
        }
    }
    uavm_pairbatch_destroy(ctx, pb);
    uavm_featureset_destroy(ctx, fs);
    if (rc != UAVM_OK) return UAVM_EFAIL;
    if (pairs_out && !matches.empty()) {                               // what the reference writes to matchPairs.match (:4492)
        *pairs_out = (uavm_matchpointpairs*)malloc(matches.size() * sizeof(uavm_matchpointpairs));
        if (!*pairs_out) return UAVM_EFAIL;
        memcpy(*pairs_out, matches.data(), matches.size() * sizeof(uavm_matchpointpairs));
        if (n_pairs_out) *n_pairs_out = (int)matches.size();
    }
    return mosaic_tail(ctx, images, n_images, matches, P, scale, result, num_mosaiced, transforms_out);
}

extern "C" int uavm_mosaic_images(uavm_ctx* ctx, const uavm_image* images, int n_images,
                                  const float* const* desc, const float* const* kp_xy, const int32_t* n_kp,
                                  const uavm_param* param_in, float scale,
                                  uavm_image* result, int* num_mosaiced, uavm_imagetransform* transforms_out)
{
    return uavm_mosaic_images_ex(ctx, images, n_images, desc, kp_xy, n_kp, param_in, scale, result, num_mosaiced, transforms_out, nullptr, nullptr);
}

// MosaicVavImages' own signature (M/MosaicWithoutPos.h:638-645): images in, mosaic out.  Features come from the GPU SIFT
// (csrc/sift.cu) with the reference's parameters SIFT(2000, 3, 0.01, 20) (M/MosaicWithoutPos.cpp:4852), one image after the other
// like SiftExtraction_Thread (:4832-4887) but without the round trip through keypoint_%d.key / discriptor_%d.xml.
extern "C" int uavm_mosaic_images_sift(uavm_ctx* ctx, const uavm_image* images, int n_images, const uavm_param* param_in, float scale,
                                       uavm_image* result, int* num_mosaiced, uavm_imagetransform* transforms_out,
                                       uavm_matchpointpairs** pairs_out, int* n_pairs_out)
{
    if (!ctx || !images || n_images < 2 || !result) return UAVM_EINVAL;
    memset(result, 0, sizeof(*result));
    if (check_images(images, n_images) != UAVM_OK) return UAVM_EINVAL;
    const int w = images[0].width, h = images[0].height;
    uavm_sift* sift = nullptr;
    int rc = uavm_sift_create(ctx, w, h, 2000, 3, 0.01, 20.0, 1.6, &sift);
    if (rc != UAVM_OK) return UAVM_EFAIL;
    const int cap = 1 << 16;
    std::vector<std::vector<float>> desc(n_images), kp(n_images);
    std::vector<int32_t> n_kp(n_images, 0);
    std::vector<uavm_keypoint> kbuf(cap);
    std::vector<float> dbuf((size_t)cap * 128);
    for (int i = 0; rc == UAVM_OK && i < n_images; i++) {
        int n = 0;
        rc = uavm_sift_detect_and_compute(ctx, sift, images[i].imageData, images[i].widthStep, 0, kbuf.data(), dbuf.data(), cap, &n);
        if (rc != UAVM_OK) break;
        n_kp[i] = n;
        desc[i].assign(dbuf.begin(), dbuf.begin() + (size_t)n * 128);
        kp[i].resize((size_t)n * 2);
        for (int k = 0; k < n; k++) { kp[i][2 * k] = kbuf[k].x; kp[i][2 * k + 1] = kbuf[k].y; }
    }
    uavm_sift_destroy(ctx, sift);
    if (rc != UAVM_OK) return UAVM_EFAIL;
    std::vector<const float*> dp(n_images), kpp(n_images);
    for (int i = 0; i < n_images; i++) { dp[i] = desc[i].data(); kpp[i] = kp[i].data(); }
    return uavm_mosaic_images_ex(ctx, images, n_images, dp.data(), kpp.data(), n_kp.data(), param_in, scale, result, num_mosaiced, transforms_out,
                                 pairs_out, n_pairs_out);
}

extern "C" int uavm_mosaic_from_matches(uavm_ctx* ctx, const uavm_image* images, int n_images,
                                        const uavm_matchpointpairs* pairs, int n_pairs, const uavm_param* param_in, float scale,
                                        uavm_image* result, int* num_mosaiced, uavm_imagetransform* transforms_out)
{
    if (!ctx || !images || n_images < 2 || !pairs || n_pairs < 0 || !result) return UAVM_EINVAL;
    memset(result, 0, sizeof(*result));
    uavm_param P;
    if (param_in) P = *param_in; else uavm_param_default(&P);
    if (check_images(images, n_images) != UAVM_OK) return UAVM_EINVAL;
    for (int m = 0; m < n_pairs; m++)
        if (pairs[m].ptA_i < 0 || pairs[m].ptA_i >= n_images || pairs[m].ptB_i < 0 || pairs[m].ptB_i >= n_images) return UAVM_EINVAL;
    if (num_mosaiced) *num_mosaiced = 0;
    std::vector<uavm_matchpointpairs> matches(pairs, pairs + n_pairs);
    return mosaic_tail(ctx, images, n_images, matches, P, scale, result, num_mosaiced, transforms_out);
}

// Stages [B] and [C] of MosaicWithoutPose on a match list (M/MosaicWithoutPos.cpp:4501-4652): largest connected component of the
// pair graph, matches touching unconnected images dropped (same removal order as :4512-4523), reference image 0 and its points
// fixed (:4525-4556), BundleAdjustmentSparse (:4591), unconnected images flagged with m[8] = 0 (:4646-4652).  `pairs` is
// modified in place (compacted, fixed flags set: exactly what the reference writes to matchPairs.txt); *n_used_out = matches
// kept.  label_out (optional): 1 for images of the largest component.  Returns -2 when nothing was accepted, the reference
// image is not in the largest component, or the normal equations are singular.
extern "C" int uavm_global_align(uavm_matchpointpairs* pairs, int n_pairs, int n_images, uavm_imagetransform* transforms_out,
                                 int32_t* label_out, int* n_used_out)
{
    if (n_images < 1 || n_pairs < 0 || (n_pairs > 0 && !pairs) || !transforms_out) return UAVM_EINVAL;
    std::vector<int32_t> label(n_images, 0);
    if (n_pairs > 0) { int rc = uavm_connected_images(pairs, n_pairs, n_images, label.data()); if (rc != UAVM_OK) return rc; }
    int n = n_pairs;
    for (int m = 0; m < n;) {                                          // same removal order as the reference (:4512-4523)
        if (label[pairs[m].ptA_i] == 0 || label[pairs[m].ptB_i] == 0) { pairs[m] = pairs[n - 1]; n--; }
        else m++;
    }
    const int ref = 0;
    std::vector<uavm_imagetransform> init(n_images);
    for (int i = 0; i < n_images; i++) {
        memset(&init[i], 0, sizeof(init[i]));
        init[i].h.m[0] = init[i].h.m[4] = init[i].h.m[8] = 1.0f;
    }
    for (int m = 0; m < n; m++) {
        if (pairs[m].ptA_i == ref) pairs[m].ptA_Fixed = 1;
        if (pairs[m].ptB_i == ref) pairs[m].ptB_Fixed = 1;
    }
    int n_fixed = 0;
    for (int i = 0; i < n_images; i++) if (label[i] == 0) { init[i].fixed = 1; n_fixed++; }
    if (label_out) memcpy(label_out, label.data(), sizeof(int32_t) * n_images);
    if (n_used_out) *n_used_out = n;
    if (init[ref].fixed == 0) { init[ref].fixed = 1; n_fixed++; }
    else if (n > 0) return UAVM_EFAIL;                                 // reference image 0 is not in the largest connected component
    if (n == 0) return UAVM_EFAIL;                                     // no image pair was accepted
    int rc = uavm_align_affine(pairs, n, init.data(), n_images, n_fixed, transforms_out);
    if (rc != UAVM_OK) return UAVM_EFAIL;
    for (int i = 0; i < n_images; i++) if (label[i] == 0) transforms_out[i].h.m[8] = 0;      // "skip me" (:4646-4652)
    return UAVM_OK;
}

// stages [B]-[D] of MosaicWithoutPose, shared by the matching path and the loadMatchPairs path
static int mosaic_tail(uavm_ctx* ctx, const uavm_image* images, int n_images, std::vector<uavm_matchpointpairs>& matches, const uavm_param& P,
                       float scale, uavm_image* result, int* num_mosaiced, uavm_imagetransform* transforms_out)
{
    const int w = images[0].width, h = images[0].height;
    uavm_canvas* cv = nullptr;
    int rc = UAVM_OK;

    // ---- [B] + [C]: connectivity, reference image, global affine alignment ----
    std::vector<uavm_imagetransform> refined(n_images);
    {
        std::vector<uavm_matchpointpairs> work(matches);
        int n_used = 0;
        rc = uavm_global_align(work.data(), (int)work.size(), n_images, refined.data(), nullptr, &n_used);
        if (num_mosaiced) *num_mosaiced = n_images;                    // numMosaiced = nImages (:4621)
        if (rc != UAVM_OK) {
            UAVM_SET_ERR(ctx, matches.empty() ? "no image pair was accepted" : "global alignment failed (%d): reference image 0 outside the largest component or singular system", rc);
            return UAVM_EFAIL;
        }
    }
    if (transforms_out) memcpy(transforms_out, refined.data(), sizeof(uavm_imagetransform) * n_images);

    // ---- [D] warp + blend: MergeImagesRefined -> LaplacianPyramidBlending(band = 5, scale) (:4663, :2180) ----
    std::vector<float> H((size_t)n_images * 9);
    for (int i = 0; i < n_images; i++) {
        memcpy(&H[(size_t)i * 9], refined[i].h.m, 36);
        for (int j = 0; j < 6; j++) H[(size_t)i * 9 + j] *= scale;                         // pImgT[i].m[j] *= scale (M/MosaicImage.cpp:2216-2223)
    }
    std::vector<int32_t> keep(n_images, 1);
    if (P.blending == 2) uavm_resample_by_overlap(H.data(), n_images, w, h, P.overlapT, keep.data());
    rc = uavm_canvas_create(ctx, n_images, w, h, H.data(), keep.data(), &cv);
    for (int i = 0; rc == UAVM_OK && i < n_images; i++)
        if (keep[i] && H[(size_t)i * 9 + 8] != 0) rc = uavm_canvas_set_image(ctx, cv, i, images[i].imageData, images[i].widthStep, 0);
    if (rc == UAVM_OK) {
        if (P.blending == 2) {
            // K6 first (it needs no pixels), then K5 only where the blend reads, then K7
            rc = uavm_canvas_seam_masks(ctx, cv);
            if (rc == UAVM_OK) rc = P.numBands <= 5 ? uavm_canvas_warp_for_blend(ctx, cv) : uavm_canvas_warp(ctx, cv);
            if (rc == UAVM_OK && P.numBands > 5) rc = uavm_canvas_seam_masks(ctx, cv);
            if (rc == UAVM_OK) rc = uavm_canvas_blend(ctx, cv, P.numBands);
        } else rc = uavm_canvas_paste(ctx, cv);
    }
    if (rc == UAVM_OK) {
        int rw = 0, rh = 0;
        uavm_canvas_result_size(cv, &rw, &rh);
        result->width = rw; result->height = rh; result->nChannels = 3;
        result->widthStep = (rw * 3 + 3) & ~3;                                             // IplImage row alignment
        result->imageData = (uint8_t*)calloc((size_t)result->widthStep * result->height, 1);
        if (!result->imageData) rc = UAVM_EFAIL;
        else rc = uavm_canvas_get_result(ctx, cv, result->imageData, result->widthStep, nullptr, 0);
        if (rc != UAVM_OK) { free(result->imageData); memset(result, 0, sizeof(*result)); }
    }
    uavm_canvas_destroy(ctx, cv);
    return rc == UAVM_OK ? UAVM_OK : UAVM_EFAIL;                        // -2: mosaic failed (:4675-4676)
}


// Chunked mosaicking of a frame sequence: the loop of MosaicUavVideo (M/MosaicWithoutPos.cpp:10252-10300) without the video
// decoding (ConvertVideo2Bmp / cvLoadImage are upstream of this library): frames [n, min(n + max_once, total))
// are mosaicked independently, the next chunk starts at n + numMosaiced.  A chunk that fails yields an empty result
// (width = height = 0, imageData = NULL) and the loop goes on, as the reference does (it ignores nRet); unlike the reference,
// a chunk that reports numMosaiced = 0 advances by the chunk length instead of looping forever.
extern "C" int uavm_mosaic_sequence(uavm_ctx* ctx, const uavm_image* images, int n_images,
                                    const float* const* desc, const float* const* kp_xy, const int32_t* n_kp,
                                    const uavm_param* param, float scale, int max_once_mosaic_num,
                                    uavm_image** results_out, int32_t** first_frame_out, int* n_results_out)
{
    if (!ctx || !images || n_images <= 1 || !desc || !kp_xy || !n_kp || !results_out || !n_results_out || max_once_mosaic_num < 2) return UAVM_EINVAL;
    *results_out = nullptr; *n_results_out = 0;
    if (first_frame_out) *first_frame_out = nullptr;
    std::vector<uavm_image> res; std::vector<int32_t> first;
    for (int n = 0; n < n_images;) {
        const int beg = n, end = (n + max_once_mosaic_num - 1 < n_images - 1) ? n + max_once_mosaic_num - 1 : n_images - 1;
        const int cnt = end - beg + 1;
        uavm_image r; memset(&r, 0, sizeof(r));
        int num = 0;
        if (cnt >= 2) uavm_mosaic_images(ctx, images + beg, cnt, desc + beg, kp_xy + beg, n_kp + beg, param, scale, &r, &num, nullptr);
        res.push_back(r); first.push_back(beg);
        n = beg + (num > 0 ? num : cnt);
    }
    *results_out = (uavm_image*)malloc(res.size() * sizeof(uavm_image));
    if (!*results_out) { for (auto& r : res) free(r.imageData); return UAVM_EFAIL; }
    memcpy(*results_out, res.data(), res.size() * sizeof(uavm_image));
    if (first_frame_out) {
        *first_frame_out = (int32_t*)malloc(first.size() * sizeof(int32_t));
        if (*first_frame_out) memcpy(*first_frame_out, first.data(), first.size() * sizeof(int32_t));
    }
    *n_results_out = (int)res.size();
    return UAVM_OK;
}
