// internal.h — shared host-side structures of libuavmosaic (not part of the C ABI).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include "../../include/uavm.h"

struct uavm_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;     // stream every launch of this ctx goes to (main or, between fork/unfork, side)
    bool own_stream = false;
    cudaStream_t main_stream = nullptr, side_stream = nullptr;
    cudaStream_t copy_stream = nullptr; // host->device frame copies (overlap with compute on `stream`, ordered by events)
    cudaEvent_t ev_fork = nullptr, ev_side = nullptr;
    bool forked = false, side_pending = false;
    int64_t launches = 0;
    bool k5_attr_set = false;          // per-device function attributes applied through THIS context (attributes are per device,
    bool k2_attr_set = false, k4_fin_attr_set = false;   // contexts are per device: a process-wide flag would skip the opt-in on a second GPU)
    size_t k4_eval_attr_smem = 0;
    char err[512] = {0};
};

#define UAVM_SET_ERR(ctx, ...)                                  \
    do {                                                        \
        if (ctx) snprintf((ctx)->err, sizeof((ctx)->err), __VA_ARGS__); \
    } while (0)

#define UAVM_CUDA(ctx, call)                                                                     \
    do {                                                                                         \
        cudaError_t e__ = (call);                                                                \
        if (e__ != cudaSuccess) {                                                                \
            UAVM_SET_ERR(ctx, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return UAVM_EFAIL;                                                                   \
        }                                                                                        \
    } while (0)

// same, but runs `cleanup` before returning (allocation paths: an OOM must not leak what was already allocated)
#define UAVM_CUDA_OR(ctx, call, cleanup)                                                         \
    do {                                                                                         \
        cudaError_t e__ = (call);                                                                \
        if (e__ != cudaSuccess) {                                                                \
            UAVM_SET_ERR(ctx, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            cudaGetLastError();                                                                  \
            cleanup;                                                                             \
            return UAVM_EFAIL;                                                                   \
        }                                                                                        \
    } while (0)

#define UAVM_CHECK_LAUNCH(ctx)                                                             \
    do {                                                                                   \
        cudaError_t e__ = cudaGetLastError();                                              \
        if (e__ != cudaSuccess) {                                                          \
            UAVM_SET_ERR(ctx, "%s:%d kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
            return UAVM_EFAIL;                                                             \
        }                                                                                  \
        (ctx)->launches++;                                                                 \
    } while (0)

// Feature pool: all images' descriptors in ONE row-major u8 matrix [pool_rows x 128] so a single TMA
// tensor map covers every pair.  Image k owns rows [row0[k], row0[k] + cap[k]), cap a multiple of
// 256; rows past n[k] are zero with a sentinel column key, so tiles never need bounds checks.
struct uavm_featureset {
    int n_images = 0;
    std::vector<int32_t> n;      // keypoints per image
    std::vector<int32_t> row0;   // first pool row per image
    std::vector<int32_t> cap;    // padded rows per image
    int64_t pool_rows = 0;
    uint8_t* d_desc = nullptr;   // [pool_rows][128] u8
    int32_t* d_norm = nullptr;   // [pool_rows] sum of squares (query side)
    int32_t* d_ckey = nullptr;   // [pool_rows] -(32*norm + (local_row & 31)), sentinel on padding (train side)
    float* d_kp = nullptr;       // [pool_rows][2] keypoint xy
    void* d_stage = nullptr;     // staging for f32 uploads
    size_t stage_bytes = 0;
    // host uploads travel on the ctx's copy stream (FIFO with the frame copies of the canvas, so a step's descriptors are
    // not overtaken by its 36 MB frames on the DMA engine); ev_up orders the pack kernel after the copy, ev_read orders a new
    // upload after the last match kernel that read the pool
    static constexpr int kUpEvents = 8;
    cudaEvent_t ev_up[kUpEvents] = {nullptr};
    int ev_up_next = 0;
    cudaEvent_t ev_read = nullptr;
    bool pool_read_since_wait = true;          // a match kernel was launched since the copy stream last waited for the compute stream
    std::vector<uint8_t> uploaded_since_wait;  // per image: its rows were rewritten since that wait (a second rewrite must wait again)
    CUtensorMap tmap_q;          // box 256 rows x 128 B (query block)
    CUtensorMap tmap_t;          // box 128 rows x 128 B (train tile)
};

struct MatchItem {
    int32_t q_row;     // pool row of this 256-row query block
    int32_t q_valid;   // valid rows in the block
    int32_t t_row;     // pool row of the train image
    int32_t n_tiles;   // ceil(n_train / 128)
    int32_t out_off;   // offset into the pair batch's match arrays
    int32_t pad[3];
};

struct PairDesc {
    int32_t img_q, img_t;
    int32_t nq, nt;
    int32_t q_row, t_row;     // pool rows
    int32_t match_off;        // offset of this pair's matches in d_train_idx/d_d2
    uint32_t seed;
};

#define UAVM_CAND_SLOTS 768       // >= (gridX*(gridY+1)+gridX+1) * int(maxNum/9) = 16*44 = 704 for the defaults
#define UAVM_RANSAC_MAX_TUPLES_FIRST 4096   // per-pair stride of the first-pass draw-group results

struct uavm_pairbatch {
    uavm_featureset* fs = nullptr;
    int n_pairs = 0;
    std::vector<PairDesc> pairs;
    PairDesc* d_pairs = nullptr;
    // match
    int n_items = 0;
    MatchItem* d_items = nullptr;
    int64_t total_q = 0;
    int32_t* d_train_idx = nullptr;   // [total_q]
    int32_t* d_d2 = nullptr;          // [total_q]
    int max_nq = 0;
    // select
    float* d_cand_xy1 = nullptr;      // [n_pairs][SLOTS][2]
    float* d_cand_xy2 = nullptr;
    int32_t* d_cand_id1 = nullptr;    // [n_pairs][SLOTS]
    int32_t* d_cand_id2 = nullptr;
    int32_t* d_cand_n = nullptr;      // [n_pairs]
    // ransac
    int tuples_first_pass = 0;
    uint32_t* d_tuple_res = nullptr;  // [n_pairs][MAX_TUPLES_FIRST] packed (valid, rejected, support)
    float* d_tuple_h = nullptr;       // [n_pairs][MAX_TUPLES_FIRST][9] hypotheses of the accepted first-pass tuples
    uint8_t* d_inlier = nullptr;      // [n_pairs][SLOTS]
    uavm_ransac_result* d_res = nullptr;  // [n_pairs]
    bool matched = false, selected = false, ransacked = false;
};

int uavm_encode_tmap_2d(uavm_ctx* ctx, CUtensorMap* tm, int dtype, void* base, uint64_t dim0, uint64_t dim1, uint64_t stride1_bytes,
                        uint32_t box0, uint32_t box1, int swizzle128);
int uavm_launch_match(uavm_ctx* ctx, uavm_pairbatch* pb);
int uavm_launch_select(uavm_ctx* ctx, uavm_pairbatch* pb, int width, int height, int gx, int gy, int max_num, double frac);
int uavm_launch_ransac(uavm_ctx* ctx, uavm_pairbatch* pb, float dist, int sample_times);
