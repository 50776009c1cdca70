// decode.cu — f4: JPEG frames decoded on the device, straight into the canvas' source pool.
// Replaces the cvLoadImage calls of the callers (_tmain, M/mosaicing.cpp:51-100; MosaicUavVideo, M/MosaicWithoutPos.cpp:10224-10308):
// the reference decodes every frame with libjpeg on one host core and then hands 36 MB of pixels over; here the caller hands
// over the JPEG bytes and the pixels are born in HBM.
// The decoder is NVIDIA's nvJPEG (library code, like cuBLAS would be for a plain GEMM — JPEG entropy decoding is not a kernel of
// this project); it is loaded at run time (libnvjpeg.so.12), so the library itself has no link-time dependency on it.  In a BGR
// pool (frame width % 16 == 0) nvJPEG writes interleaved BGR directly into the frame's slot: no staging, no conversion.
// Decoded pixels are the DECODER's: libjpeg (the reference), libjpeg-turbo (cv2) and nvJPEG use different IDCT / upsampling
// arithmetic and differ by a few grey levels; everything downstream is bit-exact on whatever the decoder produced.
#include <dlfcn.h>
#include <nvjpeg.h>
#include <string.h>
#include <thread>
#include <vector>
#include "canvas.h"

namespace {

struct NvjpegApi {
    void* lib = nullptr;
    decltype(&nvjpegCreateEx) CreateEx = nullptr;
    decltype(&nvjpegDestroy) Destroy = nullptr;
    decltype(&nvjpegJpegStateCreate) StateCreate = nullptr;
    decltype(&nvjpegJpegStateDestroy) StateDestroy = nullptr;
    decltype(&nvjpegGetImageInfo) GetImageInfo = nullptr;
    decltype(&nvjpegDecode) Decode = nullptr;
    decltype(&nvjpegDecodeBatchedInitialize) BatchedInit = nullptr;      // optional (batched entry point only)
    decltype(&nvjpegDecodeBatched) DecodeBatched = nullptr;
    decltype(&nvjpegGetHardwareDecoderInfo) HwInfo = nullptr;
};
NvjpegApi g_nj;

const char* load_nvjpeg()
{
    if (g_nj.lib) return nullptr;
    void* h = dlopen("libnvjpeg.so.12", RTLD_NOW | RTLD_LOCAL);
    if (!h) h = dlopen("libnvjpeg.so", RTLD_NOW | RTLD_LOCAL);
    if (!h) h = dlopen("/usr/local/cuda/lib64/libnvjpeg.so.12", RTLD_NOW | RTLD_LOCAL);
    if (!h) return "libnvjpeg.so.12 not found (JPEG entry points need nvJPEG)";
#define UAVM_SYM(field, name) g_nj.field = (decltype(g_nj.field))dlsym(h, name); if (!g_nj.field) return "nvJPEG symbol " name " missing";
    UAVM_SYM(CreateEx, "nvjpegCreateEx") UAVM_SYM(Destroy, "nvjpegDestroy") UAVM_SYM(StateCreate, "nvjpegJpegStateCreate")
    UAVM_SYM(StateDestroy, "nvjpegJpegStateDestroy") UAVM_SYM(GetImageInfo, "nvjpegGetImageInfo") UAVM_SYM(Decode, "nvjpegDecode")
#undef UAVM_SYM
    g_nj.BatchedInit = (decltype(g_nj.BatchedInit))dlsym(h, "nvjpegDecodeBatchedInitialize");
    g_nj.DecodeBatched = (decltype(g_nj.DecodeBatched))dlsym(h, "nvjpegDecodeBatched");
    g_nj.HwInfo = (decltype(g_nj.HwInfo))dlsym(h, "nvjpegGetHardwareDecoderInfo");
    g_nj.lib = h;
    return nullptr;
}

}  // namespace

// One decoder lane: nvJPEG states are not re-entrant, and a state's staging buffers are still being read by the device part
// of the previous decode while the host part of the next one fills them — so a lane alternates between two states and waits
// for a state's last decode (event) before reusing it.
struct JpegLane {
    nvjpegJpegState_t state[2] = {nullptr, nullptr};
    cudaEvent_t done[2] = {nullptr, nullptr};
    bool used[2] = {false, false};
    int turn = 0;
    cudaStream_t stream = nullptr;            // lanes >= 1: the worker's own stream (lane 0 decodes on the context's stream)
};

struct uavm_jpeg {
    nvjpegHandle_t handle = nullptr;
    std::vector<JpegLane> lanes;              // lane 0 always exists; more are created by the batch entry point
    int backend = 0;
    int threads = 0;                          // host threads of a batch (0: default)
    int batch_size = 0;                       // what nvjpegDecodeBatchedInitialize was last called with
    nvjpegJpegState_t batch_state = nullptr;
    cudaEvent_t fork_ev = nullptr;
    int fancy_upsampling = 1;                 // 0 when the backend refused NVJPEG_FLAGS_UPSAMPLING_WITH_INTERPOLATION
};

static nvjpegStatus_t lane_create(uavm_jpeg* j, JpegLane* L, bool own_stream)
{
    for (int s = 0; s < 2; s++) {
        const nvjpegStatus_t st = g_nj.StateCreate(j->handle, &L->state[s]);
        if (st != NVJPEG_STATUS_SUCCESS) return st;
        if (cudaEventCreateWithFlags(&L->done[s], cudaEventDisableTiming) != cudaSuccess) return NVJPEG_STATUS_ALLOCATOR_FAILURE;
    }
    if (own_stream && cudaStreamCreateWithFlags(&L->stream, cudaStreamNonBlocking) != cudaSuccess) return NVJPEG_STATUS_ALLOCATOR_FAILURE;
    return NVJPEG_STATUS_SUCCESS;
}

static void lane_destroy(JpegLane* L)
{
    for (int s = 0; s < 2; s++) {
        if (L->done[s]) { cudaEventSynchronize(L->done[s]); cudaEventDestroy(L->done[s]); }
        if (L->state[s]) g_nj.StateDestroy(L->state[s]);
    }
    if (L->stream) { cudaStreamSynchronize(L->stream); cudaStreamDestroy(L->stream); }
}

// one frame on a lane: interleaved BGR into `dst` (pitch `step`), GPU part on `stream`
static nvjpegStatus_t lane_decode(uavm_jpeg* j, JpegLane* L, const uint8_t* jpeg, size_t n_bytes, uint8_t* d_bgr, int step, cudaStream_t stream)
{
    const int s = L->turn; L->turn ^= 1;
    if (L->used[s]) cudaEventSynchronize(L->done[s]);
    nvjpegImage_t dst; memset(&dst, 0, sizeof(dst));
    dst.channel[0] = d_bgr; dst.pitch[0] = (size_t)step;
    const nvjpegStatus_t st = g_nj.Decode(j->handle, L->state[s], jpeg, n_bytes, NVJPEG_OUTPUT_BGRI, &dst, stream);
    if (st == NVJPEG_STATUS_SUCCESS) { cudaEventRecord(L->done[s], stream); L->used[s] = true; }
    return st;
}

// backend: 0 = nvJPEG's default, 1 = hybrid (Huffman on the host), 2 = GPU hybrid (Huffman on the GPU; large frames),
// 3 = the GPU's hardware JPEG engines (baseline single-scan streams; batched entry point)
extern "C" int uavm_jpeg_create(uavm_ctx* ctx, int backend, uavm_jpeg** out)
{
    if (!ctx || !out || backend < 0 || backend > 3) return UAVM_EINVAL;
    *out = nullptr;
    if (const char* e = load_nvjpeg()) { UAVM_SET_ERR(ctx, "%s", e); return UAVM_EFAIL; }
    UAVM_CUDA(ctx, cudaSetDevice(ctx->device));
    uavm_jpeg* j = new uavm_jpeg();
    j->backend = backend;
    nvjpegStatus_t st = g_nj.CreateEx((nvjpegBackend_t)backend, nullptr, nullptr, NVJPEG_FLAGS_UPSAMPLING_WITH_INTERPOLATION, &j->handle);      // libjpeg's "fancy" chroma upsampling
    if (st != NVJPEG_STATUS_SUCCESS && backend == 3) {           // the hardware engines have their own upsampler
        j->handle = nullptr; j->fancy_upsampling = 0;
        st = g_nj.CreateEx((nvjpegBackend_t)backend, nullptr, nullptr, 0, &j->handle);
    }
    if (st == NVJPEG_STATUS_SUCCESS) { j->lanes.resize(1); st = lane_create(j, &j->lanes[0], false); }
    if (st == NVJPEG_STATUS_SUCCESS && cudaEventCreateWithFlags(&j->fork_ev, cudaEventDisableTiming) != cudaSuccess) st = NVJPEG_STATUS_ALLOCATOR_FAILURE;
    if (st != NVJPEG_STATUS_SUCCESS) {
        UAVM_SET_ERR(ctx, "nvjpeg create (backend %d) -> status %d", backend, (int)st);
        for (auto& L : j->lanes) lane_destroy(&L);
        if (j->fork_ev) cudaEventDestroy(j->fork_ev);
        if (j->handle) g_nj.Destroy(j->handle);
        delete j; return UAVM_EFAIL;
    }
    *out = j;
    return UAVM_OK;
}

extern "C" void uavm_jpeg_destroy(uavm_ctx* ctx, uavm_jpeg* j)
{
    if (!j) return;
    if (ctx) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); }
    for (auto& L : j->lanes) lane_destroy(&L);
    if (j->batch_state) g_nj.StateDestroy(j->batch_state);
    if (j->fork_ev) cudaEventDestroy(j->fork_ev);
    if (j->handle) g_nj.Destroy(j->handle);
    delete j;
}

extern "C" int uavm_jpeg_info(uavm_ctx* ctx, uavm_jpeg* j, const uint8_t* jpeg, int64_t n_bytes, int* width, int* height)
{
    if (!ctx || !j || !jpeg || n_bytes <= 0 || !width || !height) return UAVM_EINVAL;
    int nc = 0; nvjpegChromaSubsampling_t ss; int ws[NVJPEG_MAX_COMPONENT], hs[NVJPEG_MAX_COMPONENT];
    const nvjpegStatus_t st = g_nj.GetImageInfo(j->handle, jpeg, (size_t)n_bytes, &nc, &ss, ws, hs);
    if (st != NVJPEG_STATUS_SUCCESS) { UAVM_SET_ERR(ctx, "nvjpegGetImageInfo -> status %d", (int)st); return UAVM_EFAIL; }
    *width = ws[0]; *height = hs[0];
    return UAVM_OK;
}

// decode into device memory as interleaved BGR, `step` bytes per row (stream ordered on ctx's stream)
extern "C" int uavm_jpeg_decode_bgr(uavm_ctx* ctx, uavm_jpeg* j, const uint8_t* jpeg, int64_t n_bytes, uint8_t* d_bgr, int step, int width, int height)
{
    if (!ctx || !j || !jpeg || n_bytes <= 0 || !d_bgr || step < 3 * width) return UAVM_EINVAL;
    UAVM_CUDA(ctx, cudaSetDevice(ctx->device));
    int w = 0, h = 0;
    { int rc = uavm_jpeg_info(ctx, j, jpeg, n_bytes, &w, &h); if (rc != UAVM_OK) return rc; }
    if (w != width || h != height) { UAVM_SET_ERR(ctx, "jpeg is %d x %d, expected %d x %d", w, h, width, height); return UAVM_EINVAL; }
    const nvjpegStatus_t st = lane_decode(j, &j->lanes[0], jpeg, (size_t)n_bytes, d_bgr, step, ctx->stream);
    if (st != NVJPEG_STATUS_SUCCESS) { UAVM_SET_ERR(ctx, "nvjpegDecode -> status %d", (int)st); return UAVM_EFAIL; }
    return UAVM_OK;
}

// source frame `image` of a canvas from JPEG bytes.  BGR pool: decoded in place; BGRA pool: decoded into a staging slot, then
// the usual conversion.
extern "C" int uavm_canvas_set_image_jpeg(uavm_ctx* ctx, uavm_canvas* cv, uavm_jpeg* j, int image, const uint8_t* jpeg, int64_t n_bytes)
{
    if (!ctx || !cv || !j || image < 0 || image >= cv->n || !jpeg || n_bytes <= 0) return UAVM_EINVAL;
    if (cv->src_bgr) {
        uint8_t* dst8 = reinterpret_cast<uint8_t*>(cv->d_src) + (size_t)image * cv->img_h * cv->src_step_px;
        return uavm_jpeg_decode_bgr(ctx, j, jpeg, n_bytes, dst8, cv->src_step_px, cv->img_w, cv->img_h);
    }
    uint8_t* tmp = nullptr;
    UAVM_CUDA(ctx, cudaMalloc(&tmp, (size_t)cv->img_h * cv->img_w * 3));
    int rc = uavm_jpeg_decode_bgr(ctx, j, jpeg, n_bytes, tmp, cv->img_w * 3, cv->img_w, cv->img_h);
    if (rc == UAVM_OK) rc = uavm_canvas_set_image(ctx, cv, image, tmp, cv->img_w * 3, 1);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(tmp);
    return rc;
}

// host threads a batch is decoded with: n >= 1 lanes (one nvJPEG state pair + CUDA stream each), 0 = default (the host's
// hardware threads, at most 32), -1 = nvJPEG's own batched decoder (nvjpegDecodeBatched) on the calling thread
extern "C" int uavm_jpeg_set_threads(uavm_jpeg* j, int n)
{
    if (!j || n < -1 || n > 64) return UAVM_EINVAL;
    j->threads = n;
    return UAVM_OK;
}

// `count` consecutive source frames from JPEG bytes (BGR pools).  The entropy (Huffman) stage of JPEG is sequential per frame
// and runs on the host in nvJPEG's hybrid backends (~18 ms for a 12 Mpx frame): the frames are spread over host threads, each
// with its own decoder lane and stream, so a batch costs count / threads frames of latency instead of count.  Stream ordered:
// the lanes start after the work already queued on ctx's stream and ctx's stream continues after the last lane.
extern "C" int uavm_canvas_set_images_jpeg(uavm_ctx* ctx, uavm_canvas* cv, uavm_jpeg* j, int first, int count, const uint8_t* const* jpegs, const int64_t* n_bytes)
{
    if (!ctx || !cv || !j || first < 0 || count <= 0 || first + count > cv->n || !jpegs || !n_bytes) return UAVM_EINVAL;
    if (!cv->src_bgr) {                                                   // no direct target: frame after frame through the conversion
        for (int k = 0; k < count; k++) { const int rc = uavm_canvas_set_image_jpeg(ctx, cv, j, first + k, jpegs[k], n_bytes[k]); if (rc != UAVM_OK) return rc; }
        return UAVM_OK;
    }
    UAVM_CUDA(ctx, cudaSetDevice(ctx->device));
    std::vector<uint8_t*> dst(count);
    for (int k = 0; k < count; k++) {
        if (!jpegs[k] || n_bytes[k] <= 0) return UAVM_EINVAL;
        int w = 0, h = 0;
        { const int rc = uavm_jpeg_info(ctx, j, jpegs[k], n_bytes[k], &w, &h); if (rc != UAVM_OK) return rc; }
        if (w != cv->img_w || h != cv->img_h) { UAVM_SET_ERR(ctx, "jpeg %d is %d x %d, expected %d x %d", first + k, w, h, cv->img_w, cv->img_h); return UAVM_EINVAL; }
        dst[k] = reinterpret_cast<uint8_t*>(cv->d_src) + (size_t)(first + k) * cv->img_h * cv->src_step_px;
    }
    if (j->threads == -1) {
        if (!g_nj.BatchedInit || !g_nj.DecodeBatched) { UAVM_SET_ERR(ctx, "this nvJPEG has no batched decoder"); return UAVM_EFAIL; }
        std::vector<size_t> len(count); std::vector<nvjpegImage_t> img(count);
        for (int k = 0; k < count; k++) {
            len[k] = (size_t)n_bytes[k]; memset(&img[k], 0, sizeof(nvjpegImage_t));
            img[k].channel[0] = dst[k]; img[k].pitch[0] = (size_t)cv->src_step_px;
        }
        nvjpegStatus_t st = NVJPEG_STATUS_SUCCESS;
        if (!j->batch_state) st = g_nj.StateCreate(j->handle, &j->batch_state);
        if (st == NVJPEG_STATUS_SUCCESS && j->batch_size != count) {
            st = g_nj.BatchedInit(j->handle, j->batch_state, count, 1, NVJPEG_OUTPUT_BGRI);
            if (st == NVJPEG_STATUS_SUCCESS) j->batch_size = count;
        }
        if (st == NVJPEG_STATUS_SUCCESS) st = g_nj.DecodeBatched(j->handle, j->batch_state, jpegs, len.data(), img.data(), ctx->stream);
        if (st != NVJPEG_STATUS_SUCCESS) { UAVM_SET_ERR(ctx, "nvjpegDecodeBatched(%d frames) -> status %d", count, (int)st); return UAVM_EFAIL; }
        return UAVM_OK;
    }
    int nt = j->threads;
    if (nt == 0) { nt = (int)std::thread::hardware_concurrency(); if (nt < 1) nt = 1; if (nt > 32) nt = 32; }
    if (nt > count) nt = count;
    while ((int)j->lanes.size() < 1 + nt) {                               // lanes 1..nt belong to the batch workers
        j->lanes.emplace_back();
        const nvjpegStatus_t st = lane_create(j, &j->lanes.back(), true);
        if (st != NVJPEG_STATUS_SUCCESS) { lane_destroy(&j->lanes.back()); j->lanes.pop_back(); UAVM_SET_ERR(ctx, "nvjpeg lane %d -> status %d", (int)j->lanes.size(), (int)st); return UAVM_EFAIL; }
    }
    UAVM_CUDA(ctx, cudaEventRecord(j->fork_ev, ctx->stream));
    std::vector<int> status(nt, 0);
    auto work = [&](int t) {
        cudaSetDevice(ctx->device);
        JpegLane* L = &j->lanes[1 + t];
        cudaStreamWaitEvent(L->stream, j->fork_ev, 0);
        for (int k = t; k < count; k += nt) {
            const nvjpegStatus_t st = lane_decode(j, L, jpegs[k], (size_t)n_bytes[k], dst[k], cv->src_step_px, L->stream);
            if (st != NVJPEG_STATUS_SUCCESS) { status[t] = (int)st; return; }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; t++) pool.emplace_back(work, t);
    work(0);
    for (auto& th : pool) th.join();
    for (int t = 0; t < nt; t++) {                                        // ctx's stream continues after every lane's last decode
        JpegLane* L = &j->lanes[1 + t];
        const int last = L->turn ^ 1;
        if (L->used[last]) UAVM_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, L->done[last], 0));
        if (L->used[last ^ 1]) UAVM_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, L->done[last ^ 1], 0));
    }
    for (int t = 0; t < nt; t++) if (status[t]) { UAVM_SET_ERR(ctx, "nvjpegDecode (lane %d) -> status %d", t, status[t]); return UAVM_EFAIL; }
    return UAVM_OK;
}

// number of hardware JPEG engines nvJPEG sees on this device (0 when there is none / the library cannot tell)
extern "C" int uavm_jpeg_hw_engines(uavm_jpeg* j)
{
    unsigned int ne = 0, nc = 0;
    if (!j || !g_nj.HwInfo || g_nj.HwInfo(j->handle, &ne, &nc) != NVJPEG_STATUS_SUCCESS) return 0;
    return (int)ne;
}

// device address of a source frame in the canvas pool (BGR pools only): lets the GPU SIFT read the frame where the decoder put it
extern "C" int uavm_canvas_image_ptr(uavm_canvas* cv, int image, const uint8_t** d_bgr, int* step)
{
    if (!cv || image < 0 || image >= cv->n || !d_bgr || !step || !cv->src_bgr) return UAVM_EINVAL;
    *d_bgr = reinterpret_cast<const uint8_t*>(cv->d_src) + (size_t)image * cv->img_h * cv->src_step_px;
    *step = cv->src_step_px;
    return UAVM_OK;
}
