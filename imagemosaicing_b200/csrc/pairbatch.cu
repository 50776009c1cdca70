// pairbatch.cu — batched pair pipeline object (match -> select -> RANSAC) and the single-pair host-buffer
// seams of the C ABI.  Mirrors the per-pair body of GetMatchedPairsOneToAllSIFTThread
// (M/MosaicWithoutPos.cpp:5083-5227) with every pair of the batch in flight at once.
#include <string.h>
#include <math.h>
#include "internal.h"

extern "C" int uavm_pairbatch_create(uavm_ctx* ctx, uavm_featureset* fs, int n_pairs, const int32_t* pair_ij, uavm_pairbatch** out)
{
    if (!ctx || !fs || !out || n_pairs <= 0 || !pair_ij) return UAVM_EINVAL;
    *out = nullptr;
    uavm_pairbatch* pb = new uavm_pairbatch();
    pb->fs = fs; pb->n_pairs = n_pairs;
    std::vector<MatchItem> items;
    int64_t off = 0;
    for (int p = 0; p < n_pairs; p++) {
        int i = pair_ij[2 * p], j = pair_ij[2 * p + 1];
        if (i < 0 || i >= fs->n_images || j < 0 || j >= fs->n_images) { delete pb; return UAVM_EINVAL; }
        PairDesc d;
        d.img_q = i; d.img_t = j; d.nq = fs->n[i]; d.nt = fs->n[j];
        d.q_row = fs->row0[i]; d.t_row = fs->row0[j];
        d.match_off = (int32_t)off; d.seed = (uint32_t)p;
        if (d.nq > pb->max_nq) pb->max_nq = d.nq;
        if (d.nt > 0) {
            for (int b = 0; b < d.nq; b += 256) {
                MatchItem it; memset(&it, 0, sizeof(it));
                it.q_row = d.q_row + b; it.q_valid = (d.nq - b) < 256 ? (d.nq - b) : 256;
                it.t_row = d.t_row; it.n_tiles = (d.nt + 127) / 128; it.out_off = (int32_t)(off + b);
                items.push_back(it);
            }
        }
        off += d.nq;
        pb->pairs.push_back(d);
    }
    if (off > 0x7fffffffLL) { delete pb; return UAVM_EINVAL; }
    pb->total_q = off; pb->n_items = (int)items.size();
    UAVM_CUDA_OR(ctx, cudaSetDevice(ctx->device), uavm_pairbatch_destroy(ctx, pb));
    size_t np = (size_t)n_pairs;
    UAVM_CUDA_OR(ctx, cudaMalloc(&pb->d_pairs, np * sizeof(PairDesc)), uavm_pairbatch_destroy(ctx, pb));
    UAVM_CUDA_OR(ctx, cudaMalloc(&pb->d_items, (items.size() + 1) * sizeof(MatchItem)), uavm_pairbatch_destroy(ctx, pb));
    UAVM_CUDA_OR(ctx, cudaMalloc(&pb->d_train_idx, (size_t)(off + 1) * 4), uavm_pairbatch_destroy(ctx, pb));
    UAVM_CUDA_OR(ctx, cudaMalloc(&pb->d_d2, (size_t)(off + 1) * 4), uavm_pairbatch_destroy(ctx, pb));
    UAVM_CUDA_OR(ctx, cudaMalloc(&pb->d_cand_xy1, np * UAVM_CAND_SLOTS * 8), uavm_pairbatch_destroy(ctx, pb));
    UAVM_CUDA_OR(ctx, cudaMalloc(&pb->d_cand_xy2, np * UAVM_CAND_SLOTS * 8), uavm_pairbatch_destroy(ctx, pb));
    UAVM_CUDA_OR(ctx, cudaMalloc(&pb->d_cand_id1, np * UAVM_CAND_SLOTS * 4), uavm_pairbatch_destroy(ctx, pb));
    UAVM_CUDA_OR(ctx, cudaMalloc(&pb->d_cand_id2, np * UAVM_CAND_SLOTS * 4), uavm_pairbatch_destroy(ctx, pb));
    UAVM_CUDA_OR(ctx, cudaMalloc(&pb->d_cand_n, np * 4), uavm_pairbatch_destroy(ctx, pb));
    UAVM_CUDA_OR(ctx, cudaMalloc(&pb->d_tuple_res, np * UAVM_RANSAC_MAX_TUPLES_FIRST * 4), uavm_pairbatch_destroy(ctx, pb));
    UAVM_CUDA_OR(ctx, cudaMalloc(&pb->d_tuple_h, np * UAVM_RANSAC_MAX_TUPLES_FIRST * 9 * 4), uavm_pairbatch_destroy(ctx, pb));
    UAVM_CUDA_OR(ctx, cudaMalloc(&pb->d_inlier, np * UAVM_CAND_SLOTS), uavm_pairbatch_destroy(ctx, pb));
    UAVM_CUDA_OR(ctx, cudaMalloc(&pb->d_res, np * sizeof(uavm_ransac_result)), uavm_pairbatch_destroy(ctx, pb));
    UAVM_CUDA_OR(ctx, cudaMemcpyAsync(pb->d_pairs, pb->pairs.data(), np * sizeof(PairDesc), cudaMemcpyHostToDevice, ctx->stream), uavm_pairbatch_destroy(ctx, pb));
    if (!items.empty())
        UAVM_CUDA_OR(ctx, cudaMemcpyAsync(pb->d_items, items.data(), items.size() * sizeof(MatchItem), cudaMemcpyHostToDevice, ctx->stream), uavm_pairbatch_destroy(ctx, pb));
    // pairs with an empty train image have no match: trainIdx = -1
    UAVM_CUDA_OR(ctx, cudaMemsetAsync(pb->d_train_idx, 0xff, (size_t)(off + 1) * 4, ctx->stream), uavm_pairbatch_destroy(ctx, pb));
    UAVM_CUDA_OR(ctx, cudaMemsetAsync(pb->d_d2, 0, (size_t)(off + 1) * 4, ctx->stream), uavm_pairbatch_destroy(ctx, pb));
    UAVM_CUDA_OR(ctx, cudaMemsetAsync(pb->d_cand_n, 0, np * 4, ctx->stream), uavm_pairbatch_destroy(ctx, pb));
    UAVM_CUDA_OR(ctx, cudaStreamSynchronize(ctx->stream), uavm_pairbatch_destroy(ctx, pb));     // host vectors go out of scope
    *out = pb;
    return UAVM_OK;
}

extern "C" void uavm_pairbatch_destroy(uavm_ctx* ctx, uavm_pairbatch* pb)
{
    if (!pb) return;
    if (ctx) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); }
    cudaFree(pb->d_pairs); cudaFree(pb->d_items); cudaFree(pb->d_train_idx); cudaFree(pb->d_d2);
    cudaFree(pb->d_cand_xy1); cudaFree(pb->d_cand_xy2); cudaFree(pb->d_cand_id1); cudaFree(pb->d_cand_id2);
    cudaFree(pb->d_cand_n); cudaFree(pb->d_tuple_res); cudaFree(pb->d_tuple_h); cudaFree(pb->d_inlier); cudaFree(pb->d_res);
    delete pb;
}

extern "C" int uavm_pairbatch_match(uavm_ctx* ctx, uavm_pairbatch* pb)
{
    if (!ctx || !pb) return UAVM_EINVAL;
    int rc = uavm_launch_match(ctx, pb);
    if (rc == UAVM_OK) pb->matched = true;
    return rc;
}
extern "C" int uavm_pairbatch_select(uavm_ctx* ctx, uavm_pairbatch* pb, int width, int height, int grid_x, int grid_y, int max_num, double frac)
{
    if (!ctx || !pb) return UAVM_EINVAL;
    if (!pb->matched) { UAVM_SET_ERR(ctx, "select before match"); return UAVM_EINVAL; }
    int rc = uavm_launch_select(ctx, pb, width, height, grid_x, grid_y, max_num, frac);
    if (rc == UAVM_OK) pb->selected = true;
    return rc;
}
extern "C" int uavm_pairbatch_ransac(uavm_ctx* ctx, uavm_pairbatch* pb, float ransac_dist, int sample_times, const uint32_t* seeds, uint32_t base_seed)
{
    if (!ctx || !pb) return UAVM_EINVAL;
    if (!pb->selected) { UAVM_SET_ERR(ctx, "ransac before select"); return UAVM_EINVAL; }
    bool changed = false;
    for (int p = 0; p < pb->n_pairs; p++) {
        uint32_t s = seeds ? seeds[p] : base_seed + (uint32_t)p;
        if (pb->pairs[p].seed != s) { pb->pairs[p].seed = s; changed = true; }
    }
    if (changed) {
        UAVM_CUDA(ctx, cudaMemcpyAsync(pb->d_pairs, pb->pairs.data(), (size_t)pb->n_pairs * sizeof(PairDesc), cudaMemcpyHostToDevice, ctx->stream));
        UAVM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    int rc = uavm_launch_ransac(ctx, pb, ransac_dist, sample_times);
    if (rc == UAVM_OK) pb->ransacked = true;
    return rc;
}

// ------------------------------------------------------------------------------------------------
// results -> host
// ------------------------------------------------------------------------------------------------
// A caller that ran a stage on the side stream (fork ... unfork) and fetches results before uavm_ctx_join would read
// them while the kernels still run: the getters join first.
static int join_side(uavm_ctx* ctx)
{
    if (ctx->forked) { UAVM_SET_ERR(ctx, "results requested between fork and unfork"); return UAVM_EINVAL; }
    return ctx->side_pending ? uavm_ctx_join(ctx) : UAVM_OK;
}

extern "C" int uavm_pairbatch_get_matches(uavm_ctx* ctx, uavm_pairbatch* pb, int pair, uavm_dmatch* out, int cap, int* n_out)
{
    if (!ctx || !pb || pair < 0 || pair >= pb->n_pairs || !out) return UAVM_EINVAL;
    { int rcj = join_side(ctx); if (rcj != UAVM_OK) return rcj; }
    const PairDesc& d = pb->pairs[pair];
    if (cap < d.nq) return UAVM_EINVAL;
    std::vector<int32_t> ti(d.nq > 0 ? d.nq : 1), dd(d.nq > 0 ? d.nq : 1);
    if (d.nq > 0) {
        UAVM_CUDA(ctx, cudaMemcpyAsync(ti.data(), pb->d_train_idx + d.match_off, (size_t)d.nq * 4, cudaMemcpyDeviceToHost, ctx->stream));
        UAVM_CUDA(ctx, cudaMemcpyAsync(dd.data(), pb->d_d2 + d.match_off, (size_t)d.nq * 4, cudaMemcpyDeviceToHost, ctx->stream));
    }
    UAVM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < d.nq; i++) {
        out[i].queryIdx = i; out[i].trainIdx = ti[i]; out[i].imgIdx = 0;
        out[i].distance = sqrtf((float)dd[i]);     // DMatch.distance = sqrt of the (exact integer) squared distance
    }
    if (n_out) *n_out = d.nq;
    return UAVM_OK;
}

extern "C" int uavm_pairbatch_get_candidates(uavm_ctx* ctx, uavm_pairbatch* pb, int pair, uavm_sfpoint* pts1, uavm_sfpoint* pts2, int cap, int* n_out)
{
    if (!ctx || !pb || pair < 0 || pair >= pb->n_pairs || !pts1 || !pts2 || !n_out) return UAVM_EINVAL;
    { int rcj = join_side(ctx); if (rcj != UAVM_OK) return rcj; }
    int n = 0;
    UAVM_CUDA(ctx, cudaMemcpyAsync(&n, pb->d_cand_n + pair, 4, cudaMemcpyDeviceToHost, ctx->stream));
    UAVM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (n > cap) return UAVM_EINVAL;
    std::vector<float> a(2 * (size_t)n + 2), b(2 * (size_t)n + 2);
    std::vector<int32_t> ia(n + 1), ib(n + 1);
    size_t base = (size_t)pair * UAVM_CAND_SLOTS;
    if (n > 0) {
        UAVM_CUDA(ctx, cudaMemcpyAsync(a.data(), pb->d_cand_xy1 + base * 2, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
        UAVM_CUDA(ctx, cudaMemcpyAsync(b.data(), pb->d_cand_xy2 + base * 2, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
        UAVM_CUDA(ctx, cudaMemcpyAsync(ia.data(), pb->d_cand_id1 + base, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
        UAVM_CUDA(ctx, cudaMemcpyAsync(ib.data(), pb->d_cand_id2 + base, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
        UAVM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    for (int i = 0; i < n; i++) {
        pts1[i].x = a[2 * i]; pts1[i].y = a[2 * i + 1]; pts1[i].id = ia[i];
        pts2[i].x = b[2 * i]; pts2[i].y = b[2 * i + 1]; pts2[i].id = ib[i];
    }
    *n_out = n;
    return UAVM_OK;
}

extern "C" int uavm_pairbatch_get_ransac(uavm_ctx* ctx, uavm_pairbatch* pb, int pair, uint8_t* inlier_mask, int cap, uavm_ransac_result* res)
{
    if (!ctx || !pb || pair < 0 || pair >= pb->n_pairs || !res) return UAVM_EINVAL;
    { int rcj = join_side(ctx); if (rcj != UAVM_OK) return rcj; }
    if (!pb->ransacked) { UAVM_SET_ERR(ctx, "get_ransac before ransac"); return UAVM_EINVAL; }
    int n = 0;
    UAVM_CUDA(ctx, cudaMemcpyAsync(&n, pb->d_cand_n + pair, 4, cudaMemcpyDeviceToHost, ctx->stream));
    UAVM_CUDA(ctx, cudaMemcpyAsync(res, pb->d_res + pair, sizeof(uavm_ransac_result), cudaMemcpyDeviceToHost, ctx->stream));
    UAVM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (inlier_mask) {
        if (cap < n) return UAVM_EINVAL;
        if (n > 0) {
            UAVM_CUDA(ctx, cudaMemcpyAsync(inlier_mask, pb->d_inlier + (size_t)pair * UAVM_CAND_SLOTS, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
            UAVM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        }
    }
    return UAVM_OK;
}

// accept rule + MatchPointPairs assembly (M/MosaicWithoutPos.cpp:5201-5221)
extern "C" int uavm_pairbatch_collect(uavm_ctx* ctx, uavm_pairbatch* pb, int min_inner_points, uavm_matchpointpairs* out, int cap, int* n_out, int* n_accepted_pairs)
{
    if (!ctx || !pb || !n_out) return UAVM_EINVAL;
    { int rcj = join_side(ctx); if (rcj != UAVM_OK) return rcj; }
    if (!pb->ransacked) { UAVM_SET_ERR(ctx, "collect before ransac"); return UAVM_EINVAL; }
    size_t np = (size_t)pb->n_pairs;
    std::vector<uavm_ransac_result> res(np);
    std::vector<int32_t> cn(np);
    std::vector<uint8_t> inl(np * UAVM_CAND_SLOTS);
    std::vector<float> a(np * UAVM_CAND_SLOTS * 2), b(np * UAVM_CAND_SLOTS * 2);
    std::vector<int32_t> ia(np * UAVM_CAND_SLOTS), ib(np * UAVM_CAND_SLOTS);
    UAVM_CUDA(ctx, cudaMemcpyAsync(res.data(), pb->d_res, np * sizeof(uavm_ransac_result), cudaMemcpyDeviceToHost, ctx->stream));
    UAVM_CUDA(ctx, cudaMemcpyAsync(cn.data(), pb->d_cand_n, np * 4, cudaMemcpyDeviceToHost, ctx->stream));
    UAVM_CUDA(ctx, cudaMemcpyAsync(inl.data(), pb->d_inlier, inl.size(), cudaMemcpyDeviceToHost, ctx->stream));
    UAVM_CUDA(ctx, cudaMemcpyAsync(a.data(), pb->d_cand_xy1, a.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
    UAVM_CUDA(ctx, cudaMemcpyAsync(b.data(), pb->d_cand_xy2, b.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
    UAVM_CUDA(ctx, cudaMemcpyAsync(ia.data(), pb->d_cand_id1, ia.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
    UAVM_CUDA(ctx, cudaMemcpyAsync(ib.data(), pb->d_cand_id2, ib.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
    UAVM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    int cnt = 0, acc = 0;
    for (size_t p = 0; p < np; p++) {
        if (res[p].n_inliers <= min_inner_points) continue;          // nInnerPoints > MIN_INNER_POINTS
        acc++;
        for (int i = 0; i < cn[p]; i++) {
            size_t k = p * UAVM_CAND_SLOTS + i;
            if (!inl[k]) continue;
            if (out) {
                if (cnt >= cap) return UAVM_EINVAL;
                uavm_matchpointpairs& m = out[cnt];
                m.ptA.x = a[2 * k]; m.ptA.y = a[2 * k + 1]; m.ptA.id = ia[k]; m.ptA_i = pb->pairs[p].img_q; m.ptA_Fixed = 0;
                m.ptB.x = b[2 * k]; m.ptB.y = b[2 * k + 1]; m.ptB.id = ib[k]; m.ptB_i = pb->pairs[p].img_t; m.ptB_Fixed = 0;
            }
            cnt++;
        }
    }
    *n_out = cnt;
    if (n_accepted_pairs) *n_accepted_pairs = acc;
    return UAVM_OK;
}

// ------------------------------------------------------------------------------------------------
// single-pair host-buffer seams
// ------------------------------------------------------------------------------------------------
extern "C" int uavm_match(uavm_ctx* ctx, const float* desc1, int n1, const float* desc2, int n2, uavm_dmatch* matches)
{
    if (!ctx || n1 < 0 || n2 < 0 || (n1 > 0 && (!desc1 || !matches)) || (n2 > 0 && !desc2)) return UAVM_EINVAL;
    if (n1 == 0) return UAVM_OK;
    int32_t n[2] = {n1, n2};
    uavm_featureset* fs = nullptr; uavm_pairbatch* pb = nullptr;
    int rc = uavm_featureset_create(ctx, 2, n, &fs);
    if (rc != UAVM_OK) return rc;
    int32_t ij[2] = {0, 1};
    rc = uavm_featureset_upload_f32(ctx, fs, 0, desc1, nullptr, 0);
    if (rc == UAVM_OK) rc = uavm_featureset_upload_f32(ctx, fs, 1, desc2, nullptr, 0);
    if (rc == UAVM_OK) rc = uavm_pairbatch_create(ctx, fs, 1, ij, &pb);
    if (rc == UAVM_OK) rc = uavm_pairbatch_match(ctx, pb);
    int got = 0;
    if (rc == UAVM_OK) rc = uavm_pairbatch_get_matches(ctx, pb, 0, matches, n1, &got);
    uavm_pairbatch_destroy(ctx, pb);
    uavm_featureset_destroy(ctx, fs);
    return rc;
}

extern "C" int uavm_select(uavm_ctx* ctx, const uavm_dmatch* matches, int n_matches, const float* kp1_xy, int n1, const float* kp2_xy, int n2,
                           int width, int height, int grid_x, int grid_y, int max_num, double frac,
                           uavm_sfpoint* pts1, uavm_sfpoint* pts2, int cap, int* n_out)
{
    if (!ctx || !matches || n_matches <= 0 || n_matches != n1 || !kp1_xy || !kp2_xy || n2 <= 0 || !pts1 || !pts2 || !n_out) return UAVM_EINVAL;
    // matches must be the full 1-NN list in queryIdx order or any order: scatter by queryIdx
    std::vector<int32_t> ti(n1, -1), dd(n1, 0);
    for (int i = 0; i < n_matches; i++) {
        int q = matches[i].queryIdx;
        if (q < 0 || q >= n1 || matches[i].trainIdx < 0 || matches[i].trainIdx >= n2) return UAVM_EINVAL;
        ti[q] = matches[i].trainIdx;
        float d = matches[i].distance;
        dd[q] = (int32_t)lrintf(d * d);            // distance = sqrt(integer d2), d2 < 2^23: exact round trip
    }
    int32_t n[2] = {n1, n2};
    uavm_featureset* fs = nullptr; uavm_pairbatch* pb = nullptr;
    int rc = uavm_featureset_create(ctx, 2, n, &fs);
    if (rc != UAVM_OK) return rc;
    int32_t ij[2] = {0, 1};
    UAVM_CUDA_OR(ctx, cudaMemcpyAsync(fs->d_kp + (size_t)fs->row0[0] * 2, kp1_xy, (size_t)n1 * 8, cudaMemcpyHostToDevice, ctx->stream), uavm_featureset_destroy(ctx, fs));
    UAVM_CUDA_OR(ctx, cudaMemcpyAsync(fs->d_kp + (size_t)fs->row0[1] * 2, kp2_xy, (size_t)n2 * 8, cudaMemcpyHostToDevice, ctx->stream), uavm_featureset_destroy(ctx, fs));
    rc = uavm_pairbatch_create(ctx, fs, 1, ij, &pb);
    if (rc == UAVM_OK) {
        cudaMemcpyAsync(pb->d_train_idx, ti.data(), (size_t)n1 * 4, cudaMemcpyHostToDevice, ctx->stream);
        cudaMemcpyAsync(pb->d_d2, dd.data(), (size_t)n1 * 4, cudaMemcpyHostToDevice, ctx->stream);
        pb->matched = true;
        rc = uavm_pairbatch_select(ctx, pb, width, height, grid_x, grid_y, max_num, frac);
    }
    if (rc == UAVM_OK) rc = uavm_pairbatch_get_candidates(ctx, pb, 0, pts1, pts2, cap, n_out);
    uavm_pairbatch_destroy(ctx, pb);
    uavm_featureset_destroy(ctx, fs);
    return rc;
}

extern "C" int uavm_ransac2d(uavm_ctx* ctx, const uavm_sfpoint* pts1, const uavm_sfpoint* pts2, int n, float ransac_dist, int sample_times,
                             uint32_t seed, uavm_sfpoint* inner1, uavm_sfpoint* inner2, int cap, uavm_ransac_result* res)
{
    if (!ctx || !res || n < 0 || (n > 0 && (!pts1 || !pts2))) return UAVM_EINVAL;
    if (n > UAVM_CAND_SLOTS) { UAVM_SET_ERR(ctx, "ransac2d: at most %d candidate pairs", UAVM_CAND_SLOTS); return UAVM_EINVAL; }
    memset(res, 0, sizeof(*res)); res->best_tuple = -1;
    if (n == 0) return UAVM_OK;                     // Ransac2D returns false on empty input (:1739-1744)
    int32_t nn[2] = {1, 1};
    uavm_featureset* fs = nullptr; uavm_pairbatch* pb = nullptr;
    int rc = uavm_featureset_create(ctx, 2, nn, &fs);
    if (rc != UAVM_OK) return rc;
    int32_t ij[2] = {0, 1};
    rc = uavm_pairbatch_create(ctx, fs, 1, ij, &pb);
    std::vector<float> a(2 * (size_t)n), b(2 * (size_t)n);
    for (int i = 0; i < n; i++) { a[2 * i] = pts1[i].x; a[2 * i + 1] = pts1[i].y; b[2 * i] = pts2[i].x; b[2 * i + 1] = pts2[i].y; }
    std::vector<uint8_t> mask(n);
    if (rc == UAVM_OK) {
        cudaMemcpyAsync(pb->d_cand_xy1, a.data(), (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream);
        cudaMemcpyAsync(pb->d_cand_xy2, b.data(), (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream);
        cudaMemcpyAsync(pb->d_cand_n, &n, 4, cudaMemcpyHostToDevice, ctx->stream);
        pb->matched = pb->selected = true;
        rc = uavm_pairbatch_ransac(ctx, pb, ransac_dist, sample_times, &seed, 0);
    }
    if (rc == UAVM_OK) rc = uavm_pairbatch_get_ransac(ctx, pb, 0, mask.data(), n, res);
    if (rc == UAVM_OK && inner1 && inner2) {
        int k = 0;
        for (int i = 0; i < n; i++)
            if (mask[i]) {
                if (k >= cap) { rc = UAVM_EINVAL; break; }
                inner1[k] = pts1[i]; inner2[k] = pts2[i]; k++;
            }
    }
    uavm_pairbatch_destroy(ctx, pb);
    uavm_featureset_destroy(ctx, fs);
    return rc;
}
