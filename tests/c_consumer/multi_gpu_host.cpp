// The C++ host of INTEGRATION.md section 2c, compilable: one thread per GPU, each with its own uavm_ctx and uavm_dist — the
// analogue of the reference's worker fan-out + PushMatchPairs (M/MosaicWithoutPos.cpp:5244-5295, :10137-10145) for pairs, and of
// LaplacianPyramidBlending (M/MosaicImage.cpp:2205-2510) for the canvas.  tests/test_cpu_oracle.py compiles and links it against
// libuavmosaic.so (every call below must match include/uavm.h) and runs it: without a GPU uavm_ctx_create returns -2 and the
// program says so and exits 0; on a box with >= 2 GPUs it runs the whole sequence on synthetic data.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include "uavm.h"

struct Shared {
    int world = 0, n_images = 0, w = 0, h = 0, n_kp = 0;
    uint8_t id[UAVM_DIST_ID_BYTES];
    std::vector<std::vector<uint8_t>> desc;          // per image: n_kp x 128
    std::vector<std::vector<float>> kp;              // per image: n_kp x 2
    std::vector<std::vector<uint8_t>> frames;        // per image: h x w x 3 BGR
    std::vector<int32_t> pairs;                      // all pairs (i, j), global order
    std::vector<int> rc;
};

static void worker(Shared* S, int rank)
{
    int& rc = S->rc[rank];
    uavm_ctx* ctx = nullptr;
    if ((rc = uavm_ctx_create(rank, &ctx)) != UAVM_OK) return;
    uavm_dist* d = nullptr;
    if ((rc = uavm_dist_init(ctx, rank, S->world, S->id, (int)sizeof(S->id), &d)) != UAVM_OK) { uavm_ctx_destroy(ctx); return; }

    // ---- pairs: pair p belongs to rank p % world; no collective while they are computed ----
    std::vector<int32_t> n_kp(S->n_images, S->n_kp);
    uavm_featureset* fs = nullptr;
    rc = uavm_featureset_create(ctx, S->n_images, n_kp.data(), &fs);
    for (int i = 0; rc == UAVM_OK && i < S->n_images; i++) rc = uavm_featureset_upload_u8(ctx, fs, i, S->desc[i].data(), S->kp[i].data(), 0);
    const int n_pairs = (int)S->pairs.size() / 2;
    std::vector<int32_t> mine; std::vector<uint32_t> seeds;
    for (int p = rank; p < n_pairs; p += S->world) { mine.push_back(S->pairs[2 * p]); mine.push_back(S->pairs[2 * p + 1]); seeds.push_back(1000u + (uint32_t)p); }
    uavm_pairbatch* pb = nullptr;
    if (rc == UAVM_OK) rc = uavm_pairbatch_create(ctx, fs, (int)seeds.size(), mine.data(), &pb);
    if (rc == UAVM_OK) rc = uavm_pairbatch_match(ctx, pb);
    if (rc == UAVM_OK) rc = uavm_pairbatch_select(ctx, pb, S->w, S->h, 3, 3, 400, 0.3);
    if (rc == UAVM_OK) rc = uavm_pairbatch_ransac(ctx, pb, 2.5f, 1000, seeds.data(), 0);
    // merged MatchPointPairs list of ALL pairs, identical on every rank
    std::vector<uavm_matchpointpairs> list((size_t)n_pairs * 400);
    int n_list = 0, n_accepted = 0;
    if (rc == UAVM_OK) rc = uavm_pairbatch_allgather(ctx, d, pb, n_pairs, 30, list.data(), (int)list.size(), &n_list, &n_accepted);

    // ---- connectivity + reference image + global alignment: a small host stage (here on every rank) ----
    std::vector<uavm_imagetransform> T(S->n_images);
    std::vector<int32_t> label(S->n_images);
    int n_used = 0;
    if (rc == UAVM_OK) rc = uavm_global_align(list.data(), n_list, S->n_images, T.data(), label.data(), &n_used);
    if (rc != UAVM_OK) fprintf(stderr, "rank %d: alignment failed (%d accepted pairs, %d matches)\n", rank, n_accepted, n_list);

    // ---- canvas: one rectangle per rank (here: row bands), exact, no halo exchange; fused blend + gather to rank 0 ----
    std::vector<float> H((size_t)S->n_images * 9);
    std::vector<int32_t> keep(S->n_images);
    for (int i = 0; i < S->n_images; i++) { memcpy(&H[9 * (size_t)i], T[i].h.m, 9 * sizeof(float)); keep[i] = label[i]; }
    uavm_canvas* cv = nullptr;
    if (rc == UAVM_OK) rc = uavm_canvas_create(ctx, S->n_images, S->w, S->h, H.data(), keep.data(), &cv);
    int cw = 0, ch = 0;
    std::vector<int32_t> rects((size_t)S->world * 4);
    if (rc == UAVM_OK) {
        uavm_canvas_layout lay; std::vector<uavm_chip_layout> chips(S->n_images);
        rc = uavm_canvas_get_layout(cv, &lay, chips.data());
        cw = lay.canvas_w; ch = lay.canvas_h;
        for (int r = 0; r < S->world; r++) {
            const int y0 = (ch * r / S->world) & ~31, y1 = r + 1 == S->world ? ch : (ch * (r + 1) / S->world) & ~31;
            rects[4 * r + 0] = 0; rects[4 * r + 1] = y0; rects[4 * r + 2] = cw; rects[4 * r + 3] = y1;
        }
    }
    if (rc == UAVM_OK) rc = uavm_canvas_set_rect(ctx, cv, rects[4 * rank], rects[4 * rank + 1], rects[4 * rank + 2], rects[4 * rank + 3]);
    for (int i = 0; rc == UAVM_OK && i < S->n_images; i++)
        if (uavm_canvas_is_active(cv, i)) rc = uavm_canvas_set_image(ctx, cv, i, S->frames[i].data(), S->w * 3, 0);
    if (rc == UAVM_OK) rc = uavm_canvas_bind_root(ctx, d, cv, 0);
    if (rc == UAVM_OK) rc = uavm_canvas_seam_masks(ctx, cv);
    if (rc == UAVM_OK) rc = uavm_canvas_warp_for_blend(ctx, cv);
    if (rc == UAVM_OK) rc = uavm_canvas_blend(ctx, cv, 5);
    if (rc == UAVM_OK) rc = uavm_canvas_gather(ctx, d, cv, rects.data(), 0);
    if (rc == UAVM_OK) rc = uavm_ctx_sync(ctx);
    if (rc == UAVM_OK && rank == 0) {
        std::vector<uint8_t> mosaic((size_t)cw * ch * 3);
        rc = uavm_canvas_get_result(ctx, cv, mosaic.data(), cw * 3, nullptr, 0);
        unsigned long long sum = 0;
        for (uint8_t v : mosaic) sum += v;
        printf("multi_gpu_host: world %d, %d accepted pairs, %d matches (%d used by the alignment), mosaic %d x %d, byte sum %llu (bound root %d)\n",
               S->world, n_accepted, n_list, n_used, cw, ch, sum, uavm_canvas_bound_root(cv));
    }
    if (rc != UAVM_OK) fprintf(stderr, "rank %d: rc %d: %s\n", rank, rc, uavm_last_error(ctx));
    if (cv) uavm_canvas_destroy(ctx, cv);
    if (pb) uavm_pairbatch_destroy(ctx, pb);
    if (fs) uavm_featureset_destroy(ctx, fs);
    uavm_dist_destroy(ctx, d);
    uavm_ctx_destroy(ctx);
}

int main(int argc, char** argv)
{
    Shared S;
    S.world = argc > 1 ? atoi(argv[1]) : 2;
    {   // is there a GPU at all?  (no CPU fallback: -2)
        uavm_ctx* probe = nullptr;
        const int rc = uavm_ctx_create(0, &probe);
        if (rc != UAVM_OK) { printf("multi_gpu_host: no sm_100 device (uavm_ctx_create -> %d), nothing to run\n", rc); return 0; }
        uavm_ctx_destroy(probe);
    }
    // a synthetic strip: every frame shows the same texture shifted by 40 % of the width; keypoints on a grid whose descriptors
    // are a hash of the world position, so consecutive frames share 1024 true correspondences (frames two apart share none)
    S.n_images = 6; S.w = 640; S.h = 480; S.n_kp = 2048;
    auto hash = [](uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; };
    const int step = (int)(0.4 * S.w);
    S.desc.resize(S.n_images); S.kp.resize(S.n_images); S.frames.resize(S.n_images);
    for (int i = 0; i < S.n_images; i++) {
        S.desc[i].resize((size_t)S.n_kp * 128); S.kp[i].resize((size_t)S.n_kp * 2); S.frames[i].resize((size_t)S.w * S.h * 3);
        for (int k = 0; k < S.n_kp; k++) {
            const int gx = k % 64, gy = k / 64;                                   // 64 x 32 grid over the frame
            const int wx = i * step + gx * 8 + 3, wy = gy * 15 + 4;               // world position: the 256 px step is 32 grid columns
            S.kp[i][2 * k] = (float)(wx - i * step); S.kp[i][2 * k + 1] = (float)wy;
            for (int t = 0; t < 128; t++) S.desc[i][(size_t)k * 128 + t] = (uint8_t)(hash((uint32_t)(wx * 7919 + wy * 104729 + t * 31)) % 120);
        }
        for (int y = 0; y < S.h; y++)
            for (int x = 0; x < S.w; x++)
                for (int c = 0; c < 3; c++) S.frames[i][((size_t)y * S.w + x) * 3 + c] = (uint8_t)(hash((uint32_t)(((x + i * step) / 8) * 131 + (y / 8) * 17 + c)) & 255);
    }
    for (int i = 0; i + 1 < S.n_images; i++) { S.pairs.push_back(i); S.pairs.push_back(i + 1); }
    for (int i = 0; i + 2 < S.n_images; i++) { S.pairs.push_back(i); S.pairs.push_back(i + 2); }
    if (uavm_dist_unique_id(S.id, (int)sizeof(S.id)) != UAVM_OK) { printf("multi_gpu_host: NCCL not available\n"); return 0; }
    S.rc.assign(S.world, 0);
    std::vector<std::thread> th;
    for (int r = 0; r < S.world; r++) th.emplace_back(worker, &S, r);
    for (auto& t : th) t.join();
    for (int r = 0; r < S.world; r++) if (S.rc[r] != UAVM_OK) return 1;
    printf("multi_gpu_host ok\n");
    return 0;
}
