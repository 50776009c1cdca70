// canvas.h — internal structures of the warp / seam-mask / blend stage (not part of the C ABI).
#pragma once
#include "internal.h"

struct ChipDesc {               // one per image, device-visible
    const uchar4* src;          // BGRA source frame
    uint32_t* chip;             // BGRA words, chip_step WORDS per row (= align4(chip_w)); alpha != 0 <=> pixel inside the source
    uint8_t* mask;              // u8 mask plane, mask_step bytes per row (= align4(chip_w)): K6 seam masks (or the validity mask,
                                // expanded from alpha on demand: uavm_canvas_mask_plane)
    float* dist;                // K6 distance map, mask_step floats per row (nullptr until K6 runs)
    int32_t keep;
    int32_t chip_w, chip_h, chip_step, mask_step;
    int32_t beg_x, beg_y;
    float sx, sy;
    float inv[9];
    float quad[8];
    int32_t src_row0;           // first row of this frame in the stacked source tensor (image index * img_h)
    int32_t affine;             // inv[6] == 0 && inv[7] == 0 && inv[8] == 1: the projective divide is exact identity
    float lineA[4], lineB[4], lineC[4], lineInv[4];   // K6: quad edges as A x + B y + C = 0 and 1/sqrt(A^2+B^2)
    int32_t need_x0, need_y0, need_x1, need_y1;       // chip pixels this context has to produce (K5): the whole chip, or what a sharded
                                                      // canvas (uavm_canvas_set_rect) can read from it
};


struct uavm_canvas {
    int n = 0, img_w = 0, img_h = 0, src_step_px = 0;
    std::vector<float> H;
    uavm_canvas_layout layout;
    std::vector<uavm_chip_layout> chips;
    std::vector<ChipDesc> desc;
    std::vector<int32_t> need_base;  // per image: the need rectangle (x0, y0, x1, y1) set by create / set_rect (uavm_canvas_warp_for_blend narrows desc's copy)
    bool need_narrowed = false, in_warp_for_blend = false;
    uchar4* d_src = nullptr;         // source pool: BGRA pixels, or (src_bgr) packed BGR rows of src_step_px BYTES
    bool src_bgr = false;            // frames stay BGR in HBM (width % 16 == 0): no conversion pass, K5 reads BGR taps
    static constexpr int kWarpEvents = 32;
    cudaEvent_t ev_warp[kWarpEvents] = {nullptr};   // BGR pool: a frame is overwritten in place, so a new upload waits for the last warp that read it
    int ev_warp_next = 0;
    std::vector<int> last_warp_ev;   // per image: index into ev_warp of the last warp launch that read it (-1: none)
    uint32_t* d_chips = nullptr;     // BGRA chips
    uint8_t* d_masks = nullptr;
    float* d_dist_max = nullptr;     // [n] per-image maximum of the distance map (as uint bits)
    void* d_k6 = nullptr;            // [n] K6Chip (masks.cu): packed per-chip constants of the distance evaluation
    uint8_t** d_mask_ptr = nullptr;  // [n] mask plane of every chip
    int32_t* d_mask_step = nullptr;  // [n]
    int32_t* d_own_bbox = nullptr;   // [n][4] K6: bounding box (min x, min y, max x, max y; chip coordinates) of the pixels a chip owns
    std::vector<int32_t> own_bbox;   // host copy, fetched by K7 (valid when own_bbox_valid)
    bool own_bbox_valid = false;
    ChipDesc* d_desc = nullptr;
    CUtensorMap tmap_src;            // all source frames as one [n * img_h][src_step_px] tensor of BGRA words (K5 footprint staging)
    bool tmap_src_ok = false;
    // staging ring for host BGR frames: slot s is filled by the ctx's copy stream and drained by the BGR->BGRA
    // kernel on the compute stream, so the PCIe copy of frame k+1 overlaps the conversion / warp of frame k
    static constexpr int kStageSlots = 3;
    uint8_t* d_stage[kStageSlots] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_copied[kStageSlots] = {nullptr, nullptr, nullptr}, ev_free[kStageSlots] = {nullptr, nullptr, nullptr};
    int stage_next = 0, stage_pitch = 0;
    size_t stage_bytes = 0;
    size_t chips_bytes = 0, masks_bytes = 0;
    int max_chip_w = 0, max_chip_h = 0;
    // result canvas (K7 / paste)
    uint8_t* d_result = nullptr;     // canvas_h x canvas_w x 3
    uint8_t* d_result_mask = nullptr;
    uint8_t* peer_result = nullptr;  // the root's d_result mapped over NVLink (uavm_canvas_bind_root, non-root ranks): level 0 of the blend writes there too
    int bound_root = -1;             // root of that binding (-1: none), set on every rank
    int blend_bound_root = -1;       // binding the last blend ran under (uavm_canvas_gather: nothing left to copy)
    void* bound_dist = nullptr;      // the uavm_dist that binding belongs to (it owns the peer mapping; one bound canvas per uavm_dist)
    // canvas rectangle owned by this context/rank (multi-GPU canvas sharding): output pixels [rect_x0, rect_x1) x
    // [rect_y0, rect_y1) (even edges); default = the whole canvas.  K6 produces seam masks on the rectangle grown by
    // kShardMargin (what K7's pyramids of the rectangle can depend on), K5 the chip pixels K7 can read.
    static constexpr int kShardMargin = 192;
    int rect_x0 = 0, rect_y0 = 0, rect_x1 = 0, rect_y1 = 0;
    bool sharded = false, lines_dirty = true;
    int result_w = 0, result_h = 0;  // size of d_result (blend: canvas layout; paste: MosaicImagesRefined's own bbox)
    bool warped = false, seamed = false, blended = false;
    bool mask_plane_valid = false;   // d_masks currently holds the seam masks (seamed) or the validity masks (expanded from alpha)
    void* blend_ws = nullptr;        // opaque workspace owned by blend.cu
};

int uavm_canvas_upload_desc(uavm_ctx* ctx, uavm_canvas* cv);
int uavm_canvas_mask_plane(uavm_ctx* ctx, uavm_canvas* cv);   // makes sure d_masks is populated (validity masks if K6 has not run)
void uavm_blend_free(uavm_canvas* cv);
void uavm_dist_forget_canvas(void* dist, uavm_canvas* cv);     // canvas destroyed / rebound: the uavm_dist drops its pointer to it
int uavm_canvas_ensure_result(uavm_ctx* ctx, uavm_canvas* cv); // allocates (and zeroes) the blend's mosaic buffers for the canvas layout
