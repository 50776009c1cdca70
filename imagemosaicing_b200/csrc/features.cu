// features.cu — context, device-resident feature pool, K1 descriptor pack kernel.
//
// Replaces the reference's per-pair feature (de)serialisation (WriteSurfKeyPoints/LoadSurfKeyPoints,
// M/MosaicWithoutPos.cpp:4682-4734, re-read for every pair at :5073-5103): descriptors are packed
// once into a u8 pool in HBM and never leave it.
#include <stdlib.h>
#include <algorithm>
#include "internal.h"

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
extern "C" void uavm_param_default(uavm_param* p) {
    if (!p) return;
    p->ransacDist = 2.5f; p->blending = 2; p->loadMatchPairs = 0; p->sampleTimes = 1000;
    p->pairWindow = 182; p->minInnerPoints = 30; p->gridX = 3; p->gridY = 3; p->maxNum = 400;
    p->matchFrac = 0.3f; p->numBands = 5; p->overlapT = 0.7f; p->seed = 20160308u;
}

extern "C" int uavm_ctx_create(int device, uavm_ctx** out) {
    if (!out) return UAVM_EINVAL;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) {
        cudaGetLastError();
        return UAVM_EFAIL;   // no CPU fallback: without a CUDA device there is no context
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return UAVM_EFAIL;
    if (prop.major != 10) {
        fprintf(stderr, "uavm: device %d is sm_%d%d; this library is built for sm_100a only\n", device, prop.major, prop.minor);
        return UAVM_EFAIL;
    }
    if (cudaSetDevice(device) != cudaSuccess) return UAVM_EFAIL;
    uavm_ctx* c = new uavm_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return UAVM_EFAIL; }
    c->own_stream = true;
    c->main_stream = c->stream;
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);          // side stream = highest priority: its (latency-bound)
    const char* sp = getenv("UAVM_SIDE_PRIO");                     // experiment knob: "low" gives the side stream the lowest priority
    if (cudaStreamCreateWithPriority(&c->side_stream, cudaStreamNonBlocking, (sp && sp[0] == 'l') ? prio_lo : prio_hi) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||   // blocks are placed first

        cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_side, cudaEventDisableTiming) != cudaSuccess) { delete c; return UAVM_EFAIL; }
    *out = c;
    return UAVM_OK;
}
extern "C" void uavm_ctx_destroy(uavm_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->own_stream && ctx->main_stream) cudaStreamDestroy(ctx->main_stream);
    if (ctx->side_stream) cudaStreamDestroy(ctx->side_stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_side) cudaEventDestroy(ctx->ev_side);
    delete ctx;
}
extern "C" int uavm_ctx_set_stream(uavm_ctx* ctx, void* s) {
    if (!ctx) return UAVM_EINVAL;
    if (ctx->forked) return UAVM_EINVAL;
    if (ctx->own_stream && ctx->main_stream) { cudaStreamDestroy(ctx->main_stream); ctx->own_stream = false; }
    ctx->stream = ctx->main_stream = (cudaStream_t)s;
    return UAVM_OK;
}
extern "C" int uavm_ctx_sync(uavm_ctx* ctx) {
    if (!ctx) return UAVM_EINVAL;
    UAVM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->side_pending && !ctx->forked) { UAVM_CUDA(ctx, cudaStreamSynchronize(ctx->side_stream)); }
    return UAVM_OK;
}
// fork / unfork / join: run independent stages of one step concurrently.  After fork() every launch of this
// ctx goes to an internal side stream that first waits for all work queued so far; unfork() switches back to
// the main stream without waiting; join() makes the main stream wait for the side work.
extern "C" int uavm_ctx_fork(uavm_ctx* ctx) {
    if (!ctx || ctx->forked) return UAVM_EINVAL;
    UAVM_CUDA(ctx, cudaEventRecord(ctx->ev_fork, ctx->main_stream));
    UAVM_CUDA(ctx, cudaStreamWaitEvent(ctx->side_stream, ctx->ev_fork, 0));
    ctx->stream = ctx->side_stream; ctx->forked = true; ctx->side_pending = true;
    return UAVM_OK;
}
extern "C" int uavm_ctx_unfork(uavm_ctx* ctx) {
    if (!ctx || !ctx->forked) return UAVM_EINVAL;
    UAVM_CUDA(ctx, cudaEventRecord(ctx->ev_side, ctx->side_stream));
    ctx->stream = ctx->main_stream; ctx->forked = false;
    return UAVM_OK;
}
extern "C" int uavm_ctx_join(uavm_ctx* ctx) {
    if (!ctx || ctx->forked) return UAVM_EINVAL;
    if (ctx->side_pending) { UAVM_CUDA(ctx, cudaStreamWaitEvent(ctx->main_stream, ctx->ev_side, 0)); ctx->side_pending = false; }
    return UAVM_OK;
}
extern "C" const char* uavm_last_error(const uavm_ctx* ctx) { return ctx ? ctx->err : "null context"; }
extern "C" int64_t uavm_ctx_launch_count(const uavm_ctx* ctx) { return ctx ? ctx->launches : 0; }
extern "C" int uavm_ctx_sm_count(const uavm_ctx* ctx) { return ctx ? ctx->sm_count : 0; }
extern "C" void uavm_free(void* p) { free(p); }

// ------------------------------------------------------------------------------------------------
// K1: pack descriptors to u8, row norms and column keys.  One warp per descriptor row.
//   norm[r] = sum_k d[r][k]^2                      (query-side constant of |a-b|^2)
//   ckey[r] = -(32*norm[r] + (local_row & 31))     (train-side key, NEGATED: K2 maximises 64 a.b + ckey — a shift-add, LEA — so
//                                                   one integer max yields min distance AND, in the low 5 bits, lowest index)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) k1_pack_rows(const T* __restrict__ src, int n, uint8_t* __restrict__ dst,
                                                     int32_t* __restrict__ norm, int32_t* __restrict__ ckey) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (warp >= n) return;
    uint32_t v[4];
    if (sizeof(T) == 4) {
        float4 f = reinterpret_cast<const float4*>(src)[(size_t)warp * 32 + lane];
        float ff[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
        for (int i = 0; i < 4; i++) {
            int q = __float2int_rn(ff[i]);
            v[i] = (uint32_t)min(max(q, 0), 255);
        }
    } else {
        uchar4 u = reinterpret_cast<const uchar4*>(src)[(size_t)warp * 32 + lane];
        v[0] = u.x; v[1] = u.y; v[2] = u.z; v[3] = u.w;
    }
    reinterpret_cast<uint32_t*>(dst)[(size_t)warp * 32 + lane] = v[0] | (v[1] << 8) | (v[2] << 16) | (v[3] << 24);
    int32_t s = (int32_t)(v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3]);
    s = __reduce_add_sync(0xffffffffu, s);
    if (lane == 0) {
        norm[warp] = s;
        ckey[warp] = -(s * 32 + (warp & 31));
    }
}

__global__ void k1_fill_sentinel(int32_t* __restrict__ ckey, int32_t* __restrict__ norm, int64_t rows) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < rows) {
        ckey[i] = -(0x7fffffe0 | (int32_t)(i & 31));   // key part (>>5) above any real key (<= 128*255^2), negated like the rest
        norm[i] = 0;
    }
}

// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// generic 2-D tiled tensor map (dim0 = contiguous), shared by the descriptor pool (match) and the source frames (warp)
int uavm_encode_tmap_2d(uavm_ctx* ctx, CUtensorMap* tm, int dtype, void* base, uint64_t dim0, uint64_t dim1, uint64_t stride1_bytes,
                        uint32_t box0, uint32_t box1, int swizzle128) {
    static PFN_encodeTiled encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        UAVM_CUDA(ctx, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess) { UAVM_SET_ERR(ctx, "cuTensorMapEncodeTiled not available"); return UAVM_EFAIL; }
        encode = (PFN_encodeTiled)fn;
    }
    cuuint64_t gdim[2] = {(cuuint64_t)dim0, (cuuint64_t)dim1};
    cuuint64_t gstride[1] = {(cuuint64_t)stride1_bytes};
    cuuint32_t box[2] = {(cuuint32_t)box0, (cuuint32_t)box1};
    cuuint32_t estride[2] = {1, 1};
    CUresult r = encode(tm, (CUtensorMapDataType)dtype, 2, base, gdim, gstride, box, estride,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { UAVM_SET_ERR(ctx, "cuTensorMapEncodeTiled failed: %d", (int)r); return UAVM_EFAIL; }
    return UAVM_OK;
}

static int make_tmap(uavm_ctx* ctx, CUtensorMap* tm, void* base, int64_t rows, int box_rows) {
    return uavm_encode_tmap_2d(ctx, tm, (int)CU_TENSOR_MAP_DATA_TYPE_UINT8, base, 128, (uint64_t)rows, 128, 128, (uint32_t)box_rows, 1);
}

extern "C" int uavm_featureset_create(uavm_ctx* ctx, int n_images, const int32_t* n_keypoints, uavm_featureset** out) {
    if (!ctx || !out || n_images <= 0 || !n_keypoints) return UAVM_EINVAL;
    *out = nullptr;
    uavm_featureset* fs = new uavm_featureset();
    fs->n_images = n_images;
    int64_t rows = 0; int max_n = 0;
    for (int i = 0; i < n_images; i++) {
        if (n_keypoints[i] < 0) { delete fs; return UAVM_EINVAL; }
        int cap = ((n_keypoints[i] + 255) / 256) * 256;
        if (cap == 0) cap = 256;
        fs->n.push_back(n_keypoints[i]); fs->row0.push_back((int32_t)rows); fs->cap.push_back(cap);
        rows += cap;
        if (n_keypoints[i] > max_n) max_n = n_keypoints[i];
    }
    if (rows > 0x7fffff00LL) { delete fs; UAVM_SET_ERR(ctx, "feature pool too large"); return UAVM_EINVAL; }
    fs->pool_rows = rows;
    UAVM_CUDA_OR(ctx, cudaSetDevice(ctx->device), uavm_featureset_destroy(ctx, fs));
    UAVM_CUDA_OR(ctx, cudaMalloc(&fs->d_desc, (size_t)rows * 128), uavm_featureset_destroy(ctx, fs));
    UAVM_CUDA_OR(ctx, cudaMalloc(&fs->d_norm, (size_t)rows * 4), uavm_featureset_destroy(ctx, fs));
    UAVM_CUDA_OR(ctx, cudaMalloc(&fs->d_ckey, (size_t)rows * 4), uavm_featureset_destroy(ctx, fs));
    UAVM_CUDA_OR(ctx, cudaMalloc(&fs->d_kp, (size_t)rows * 8), uavm_featureset_destroy(ctx, fs));
    fs->stage_bytes = (size_t)(max_n > 0 ? max_n : 1) * 128 * sizeof(float);
    UAVM_CUDA_OR(ctx, cudaMalloc(&fs->d_stage, fs->stage_bytes), uavm_featureset_destroy(ctx, fs));
    UAVM_CUDA_OR(ctx, cudaMemsetAsync(fs->d_desc, 0, (size_t)rows * 128, ctx->stream), uavm_featureset_destroy(ctx, fs));
    UAVM_CUDA_OR(ctx, cudaMemsetAsync(fs->d_kp, 0, (size_t)rows * 8, ctx->stream), uavm_featureset_destroy(ctx, fs));
    k1_fill_sentinel<<<(unsigned)((rows + 255) / 256), 256, 0, ctx->stream>>>(fs->d_ckey, fs->d_norm, rows);
    UAVM_CHECK_LAUNCH(ctx);
    int rc = make_tmap(ctx, &fs->tmap_q, fs->d_desc, rows, 256);
    if (rc == UAVM_OK) rc = make_tmap(ctx, &fs->tmap_t, fs->d_desc, rows, 128);
    if (rc != UAVM_OK) { uavm_featureset_destroy(ctx, fs); return rc; }
    *out = fs;
    return UAVM_OK;
}

extern "C" void uavm_featureset_destroy(uavm_ctx* ctx, uavm_featureset* fs) {
    if (!fs) return;
    if (ctx) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); }
    if (ctx && ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
    cudaFree(fs->d_desc); cudaFree(fs->d_norm); cudaFree(fs->d_ckey); cudaFree(fs->d_kp); cudaFree(fs->d_stage);
    for (int e = 0; e < uavm_featureset::kUpEvents; e++) if (fs->ev_up[e]) cudaEventDestroy(fs->ev_up[e]);
    if (fs->ev_read) cudaEventDestroy(fs->ev_read);
    delete fs;
}

template <typename T>
static int upload_impl(uavm_ctx* ctx, uavm_featureset* fs, int image, const T* desc, const float* kp_xy, int is_device) {
    if (!ctx || !fs || image < 0 || image >= fs->n_images) return UAVM_EINVAL;
    int n = fs->n[image];
    if (n == 0) return UAVM_OK;
    if (!desc) return UAVM_EINVAL;
    size_t r0 = (size_t)fs->row0[image];
    const T* src = desc;
    UAVM_CUDA(ctx, cudaSetDevice(ctx->device));
    bool kp_done = false;
    if (!is_device) {
        if (sizeof(T) == 1) {
            // u8 host descriptors land directly in their pool rows and are "packed" in place (each warp reads its row before
            // writing it back): no staging buffer.  The copies go to the ctx's copy stream, back to back and ahead of any frame
            // copy queued later; the pack kernel on the compute stream waits for its own image only.
            if (!fs->ev_read) {
                UAVM_CUDA(ctx, cudaEventCreateWithFlags(&fs->ev_read, cudaEventDisableTiming));
                for (int e = 0; e < uavm_featureset::kUpEvents; e++) UAVM_CUDA(ctx, cudaEventCreateWithFlags(&fs->ev_up[e], cudaEventDisableTiming));
            }
            // the rows may still be read or written by work queued earlier on the compute stream (a previous match, this image's
            // previous pack): the copy stream waits for the compute stream once per round of uploads
            if (fs->uploaded_since_wait.size() != (size_t)fs->n_images) fs->uploaded_since_wait.assign(fs->n_images, 0);
            if (fs->pool_read_since_wait || fs->uploaded_since_wait[image]) {
                UAVM_CUDA(ctx, cudaEventRecord(fs->ev_read, ctx->stream));
                UAVM_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, fs->ev_read, 0));
                fs->pool_read_since_wait = false;
                std::fill(fs->uploaded_since_wait.begin(), fs->uploaded_since_wait.end(), 0);
            }
            fs->uploaded_since_wait[image] = 1;
            UAVM_CUDA(ctx, cudaMemcpyAsync(fs->d_desc + r0 * 128, desc, (size_t)n * 128, cudaMemcpyHostToDevice, ctx->copy_stream));
            if (kp_xy) { UAVM_CUDA(ctx, cudaMemcpyAsync(fs->d_kp + r0 * 2, kp_xy, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->copy_stream)); kp_done = true; }
            const int e = fs->ev_up_next; fs->ev_up_next = (e + 1) % uavm_featureset::kUpEvents;
            UAVM_CUDA(ctx, cudaEventRecord(fs->ev_up[e], ctx->copy_stream));
            UAVM_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, fs->ev_up[e], 0));
            src = (const T*)(fs->d_desc + r0 * 128);
        } else {
            // stream-ordered: the staging buffer is reused only after the previous pack kernel (same stream)
            UAVM_CUDA(ctx, cudaMemcpyAsync(fs->d_stage, desc, (size_t)n * 128 * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
            src = (const T*)fs->d_stage;
        }
    }
    k1_pack_rows<T><<<(n * 32 + 255) / 256, 256, 0, ctx->stream>>>(src, n, fs->d_desc + r0 * 128, fs->d_norm + r0, fs->d_ckey + r0);
    UAVM_CHECK_LAUNCH(ctx);
    if (kp_xy && !kp_done)
        UAVM_CUDA(ctx, cudaMemcpyAsync(fs->d_kp + r0 * 2, kp_xy, (size_t)n * 8, is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx->stream));
    return UAVM_OK;
}
extern "C" int uavm_featureset_upload_f32(uavm_ctx* ctx, uavm_featureset* fs, int image, const float* desc, const float* kp_xy, int is_device) {
    return upload_impl<float>(ctx, fs, image, desc, kp_xy, is_device);
}
extern "C" int uavm_featureset_upload_u8(uavm_ctx* ctx, uavm_featureset* fs, int image, const uint8_t* desc, const float* kp_xy, int is_device) {
    return upload_impl<uint8_t>(ctx, fs, image, desc, kp_xy, is_device);
}
