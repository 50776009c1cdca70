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
    g_nj.lib = h;
    return nullptr;
}

}  // namespace

struct uavm_jpeg {
    nvjpegHandle_t handle = nullptr;
    nvjpegJpegState_t state = nullptr;
    int backend = 0;
};

// backend: 0 = nvJPEG's default, 1 = hybrid (Huffman on the host), 2 = GPU hybrid (Huffman on the GPU; large frames)
extern "C" int uavm_jpeg_create(uavm_ctx* ctx, int backend, uavm_jpeg** out)
{
    if (!ctx || !out || backend < 0 || backend > 2) return UAVM_EINVAL;
    *out = nullptr;
    if (const char* e = load_nvjpeg()) { UAVM_SET_ERR(ctx, "%s", e); return UAVM_EFAIL; }
    UAVM_CUDA(ctx, cudaSetDevice(ctx->device));
    uavm_jpeg* j = new uavm_jpeg();
    j->backend = backend;
    nvjpegStatus_t st = g_nj.CreateEx((nvjpegBackend_t)backend, nullptr, nullptr, NVJPEG_FLAGS_UPSAMPLING_WITH_INTERPOLATION, &j->handle);      // libjpeg's "fancy" chroma upsampling
    if (st == NVJPEG_STATUS_SUCCESS) st = g_nj.StateCreate(j->handle, &j->state);
    if (st != NVJPEG_STATUS_SUCCESS) {
        UAVM_SET_ERR(ctx, "nvjpeg create (backend %d) -> status %d", backend, (int)st);
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
    if (j->state) g_nj.StateDestroy(j->state);
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
    nvjpegImage_t dst; memset(&dst, 0, sizeof(dst));
    dst.channel[0] = d_bgr; dst.pitch[0] = (size_t)step;
    const nvjpegStatus_t st = g_nj.Decode(j->handle, j->state, jpeg, (size_t)n_bytes, NVJPEG_OUTPUT_BGRI, &dst, ctx->stream);
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

// device address of a source frame in the canvas pool (BGR pools only): lets the GPU SIFT read the frame where the decoder put it
extern "C" int uavm_canvas_image_ptr(uavm_canvas* cv, int image, const uint8_t** d_bgr, int* step)
{
    if (!cv || image < 0 || image >= cv->n || !d_bgr || !step || !cv->src_bgr) return UAVM_EINVAL;
    *d_bgr = reinterpret_cast<const uint8_t*>(cv->d_src) + (size_t)image * cv->img_h * cv->src_step_px;
    *step = cv->src_step_px;
    return UAVM_OK;
}
