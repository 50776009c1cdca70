// dist.cu — multi-GPU plumbing behind the C ABI: one uavm_ctx per GPU / rank, NCCL over NVLink.
//
// The reference's analogue is the worker fan-out of GetMatchedPairsOneToAllSIFT_MultiThread
// (M/MosaicWithoutPos.cpp:5244-5295: threads striding over image pairs, results appended under a mutex,
// PushMatchPairs :10137-10145).  Here image pairs are sharded round-robin over ranks (pair p belongs to rank
// p % world), every rank runs match -> select -> RANSAC on its shard with no exchange, and ONE collective merges the
// results: uavm_pairbatch_allgather packs each pair's inliers into a fixed-size device record, all-gathers the records
// over NVLink, and compacts them ON THE DEVICE into the MatchPointPairs list in global pair order — byte-identical to
// what a single context produces with uavm_pairbatch_collect.  The canvas is sharded into rectangles
// (uavm_canvas_set_rect); uavm_canvas_gather moves the finished rectangles to the root with grouped ncclSend / ncclRecv.
//
// NCCL is loaded at run time (dlopen "libnccl.so.2"): a process that already carries an NCCL (e.g. torch's bundled one)
// shares it, a plain C++ host gets the system library, and single-GPU users need no NCCL at all.
#include <dlfcn.h>
#include <nccl.h>
#include <stdlib.h>
#include <string.h>
#include "canvas.h"

namespace {

struct NcclApi {
    void* handle = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclBroadcast) Broadcast = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
};
NcclApi g_nccl;

const char* load_nccl()
{
    if (g_nccl.handle) return nullptr;
    // an NCCL that is already in the process wins (one NCCL per process: e.g. torch's bundled libnccl.so.2, which torch's own
    // libraries were linked against); UAVM_NCCL_LIB names a specific file; otherwise the system library
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h && getenv("UAVM_NCCL_LIB")) h = dlopen(getenv("UAVM_NCCL_LIB"), RTLD_NOW | RTLD_LOCAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
    if (!h) return "libnccl.so.2 not found (multi-GPU entry points need NCCL)";
#define UAVM_SYM(field, name) g_nccl.field = (decltype(g_nccl.field))dlsym(h, name); if (!g_nccl.field) return "NCCL symbol " name " missing";
    UAVM_SYM(GetUniqueId, "ncclGetUniqueId") UAVM_SYM(CommInitRank, "ncclCommInitRank") UAVM_SYM(CommDestroy, "ncclCommDestroy")
    UAVM_SYM(AllGather, "ncclAllGather") UAVM_SYM(Send, "ncclSend") UAVM_SYM(Recv, "ncclRecv") UAVM_SYM(GroupStart, "ncclGroupStart")
    UAVM_SYM(GroupEnd, "ncclGroupEnd") UAVM_SYM(GetErrorString, "ncclGetErrorString") UAVM_SYM(Broadcast, "ncclBroadcast") UAVM_SYM(AllReduce, "ncclAllReduce")
#undef UAVM_SYM
    g_nccl.handle = h;
    return nullptr;
}

#define UAVM_NCCL(ctx, call)                                                                            \
    do {                                                                                                \
        ncclResult_t r__ = (call);                                                                      \
        if (r__ != ncclSuccess) {                                                                       \
            UAVM_SET_ERR(ctx, "%s:%d %s -> %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r__)); \
            return UAVM_EFAIL;                                                                          \
        }                                                                                               \
    } while (0)

// One record per pair: what uavm_pairbatch_collect needs of it.
constexpr int kRecCap = UAVM_CAND_SLOTS;
struct __align__(8) PairRecord {
    int32_t n_inliers;        // Ransac2D's inlier count (accept rule: > minInnerPoints)
    int32_t n_entries;        // inliers stored below, in candidate order
    int32_t img_q, img_t;
    struct Entry { float x1, y1; int32_t id1; float x2, y2; int32_t id2; } e[kRecCap];
};

// one CTA per local pair: inlier mask + candidates -> record (entries in candidate order, like the host loop of collect)
__global__ void __launch_bounds__(256)
k_dist_pack(const PairDesc* __restrict__ pairs, int n_local, const int32_t* __restrict__ cand_n, const uint8_t* __restrict__ inlier,
            const float* __restrict__ xy1, const float* __restrict__ xy2, const int32_t* __restrict__ id1, const int32_t* __restrict__ id2,
            const uavm_ransac_result* __restrict__ res, PairRecord* __restrict__ rec)
{
    const int p = blockIdx.x;
    PairRecord& R = rec[p];
    if (p >= n_local) {                                  // padding record of a rank with fewer pairs
        if (threadIdx.x == 0) { R.n_inliers = 0; R.n_entries = 0; R.img_q = -1; R.img_t = -1; }
        return;
    }
    __shared__ int wsum[8];
    __shared__ int base;
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    const int n = min(cand_n[p], kRecCap);
    const size_t o = (size_t)p * UAVM_CAND_SLOTS;
    for (int i0 = 0; i0 < n; i0 += 256) {
        const int i = i0 + threadIdx.x;
        const bool in = i < n && inlier[o + i] != 0;
        const unsigned m = __ballot_sync(0xffffffffu, in);
        if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = __popc(m);
        __syncthreads();
        int off = base;
        for (int w = 0; w < (int)(threadIdx.x >> 5); w++) off += wsum[w];
        if (in) {
            PairRecord::Entry& e = R.e[off + __popc(m & ((1u << (threadIdx.x & 31)) - 1u))];
            e.x1 = xy1[2 * (o + i)]; e.y1 = xy1[2 * (o + i) + 1]; e.id1 = id1[o + i];
            e.x2 = xy2[2 * (o + i)]; e.y2 = xy2[2 * (o + i) + 1]; e.id2 = id2[o + i];
        }
        __syncthreads();
        if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < 8; w++) t += wsum[w]; base += t; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { R.n_inliers = res[p].n_inliers; R.n_entries = base; R.img_q = pairs[p].img_q; R.img_t = pairs[p].img_t; }
}

// global pair p lives at record (p % world) * n_slot + p / world of the gathered buffer.
// offsets[p] = number of MatchPointPairs written before pair p (accepted pairs only); offsets[n_pairs] = total.
__global__ void __launch_bounds__(1024)
k_dist_offsets(const PairRecord* __restrict__ rec, int n_pairs, int world, int n_slot, int min_inner, int32_t* __restrict__ offsets, int32_t* __restrict__ n_accepted)
{
    __shared__ int wsum[32];
    __shared__ int carry, acc;
    if (threadIdx.x == 0) { carry = 0; acc = 0; }
    __syncthreads();
    for (int p0 = 0; p0 < n_pairs; p0 += 1024) {
        const int p = p0 + threadIdx.x;
        int c = 0;
        bool accepted = false;
        if (p < n_pairs) {
            const PairRecord& R = rec[(size_t)(p % world) * n_slot + p / world];
            accepted = R.n_inliers > min_inner;
            if (accepted) c = R.n_entries;
        }
        const unsigned okm = __ballot_sync(0xffffffffu, accepted);
        int v = c;                                        // inclusive warp scan
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) { const int t = __shfl_up_sync(0xffffffffu, v, s); if ((threadIdx.x & 31) >= s) v += t; }
        if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = v;
        if ((threadIdx.x & 31) == 0) atomicAdd(&acc, __popc(okm));
        __syncthreads();
        int off = carry;
        for (int w = 0; w < (int)(threadIdx.x >> 5); w++) off += wsum[w];
        if (p < n_pairs) offsets[p] = off + v - c;
        __syncthreads();
        if (threadIdx.x == 1023) carry = off + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) { offsets[n_pairs] = carry; *n_accepted = acc; }
}

__global__ void __launch_bounds__(128)
k_dist_compact(const PairRecord* __restrict__ rec, int n_pairs, int world, int n_slot, int min_inner, const int32_t* __restrict__ offsets,
               uavm_matchpointpairs* __restrict__ out)
{
    const int p = blockIdx.x;
    const PairRecord& R = rec[(size_t)(p % world) * n_slot + p / world];
    if (R.n_inliers <= min_inner) return;                 // nInnerPoints > MIN_INNER_POINTS (M/MosaicWithoutPos.cpp:5201)
    uavm_matchpointpairs* o = out + offsets[p];
    for (int i = threadIdx.x; i < R.n_entries; i += blockDim.x) {
        const PairRecord::Entry e = R.e[i];
        uavm_matchpointpairs m;
        m.ptA.x = e.x1; m.ptA.y = e.y1; m.ptA.id = e.id1; m.ptA_i = R.img_q; m.ptA_Fixed = 0;
        m.ptB.x = e.x2; m.ptB.y = e.y2; m.ptB.id = e.id2; m.ptB_i = R.img_t; m.ptB_Fixed = 0;
        o[i] = m;
    }
}

}  // namespace

struct uavm_dist {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1, device = 0;
    // pair gather workspace
    PairRecord* d_send = nullptr; PairRecord* d_recv = nullptr; size_t send_cap = 0, recv_cap = 0;
    int32_t* d_offsets = nullptr; size_t off_cap = 0;
    int32_t* d_nacc = nullptr;
    uavm_matchpointpairs* d_dense = nullptr; size_t dense_cap = 0;
    // canvas gather workspace (root)
    uint8_t* d_tmp = nullptr; size_t tmp_cap = 0;
    // direct gather: the root's result buffer mapped into this process (CUDA IPC), keyed by its handle
    uint8_t* d_ipc = nullptr;                         // [64 handle bytes | 4-byte barrier word]
    cudaIpcMemHandle_t peer_handle; void* peer_ptr = nullptr; bool peer_failed = false;
    uavm_canvas* bound_cv = nullptr;                  // the canvas whose blend writes through peer_ptr (uavm_canvas_bind_root), at most one
};

// the mapping in d->peer_ptr is about to change or go away: the canvas that writes through it must stop doing so
static void unbind_current(uavm_dist* d)
{
    if (!d->bound_cv) return;
    d->bound_cv->peer_result = nullptr; d->bound_cv->bound_root = -1; d->bound_cv->bound_dist = nullptr;
    d->bound_cv = nullptr;
}

void uavm_dist_forget_canvas(void* dist, uavm_canvas* cv)
{
    uavm_dist* d = static_cast<uavm_dist*>(dist);
    if (d && d->bound_cv == cv) d->bound_cv = nullptr;
}

extern "C" int uavm_dist_unique_id(uint8_t* id_out, int id_bytes)
{
    if (!id_out || id_bytes < (int)sizeof(ncclUniqueId)) return UAVM_EINVAL;
    memset(id_out, 0, (size_t)id_bytes);
    if (load_nccl()) return UAVM_EFAIL;
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != ncclSuccess) return UAVM_EFAIL;
    memcpy(id_out, &id, sizeof(id));
    return UAVM_OK;
}

extern "C" int uavm_dist_init(uavm_ctx* ctx, int rank, int world, const uint8_t* id, int id_bytes, uavm_dist** out)
{
    if (!ctx || !out || world < 1 || rank < 0 || rank >= world || !id || id_bytes < (int)sizeof(ncclUniqueId)) return UAVM_EINVAL;
    *out = nullptr;
    UAVM_CUDA(ctx, cudaSetDevice(ctx->device));
    uavm_dist* d = new uavm_dist();
    d->rank = rank; d->world = world; d->device = ctx->device;
    if (world == 1) { *out = d; return UAVM_OK; }          // a single rank needs no communicator (and no NCCL in the process)
    if (const char* e = load_nccl()) { UAVM_SET_ERR(ctx, "%s", e); delete d; return UAVM_EFAIL; }
    ncclUniqueId uid; memcpy(&uid, id, sizeof(uid));
    ncclResult_t r = g_nccl.CommInitRank(&d->comm, world, uid, rank);
    if (r != ncclSuccess) { UAVM_SET_ERR(ctx, "ncclCommInitRank -> %s", g_nccl.GetErrorString(r)); delete d; return UAVM_EFAIL; }
    *out = d;
    return UAVM_OK;
}

extern "C" void uavm_dist_destroy(uavm_ctx* ctx, uavm_dist* d)
{
    if (!d) return;
    if (ctx) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); }
    cudaFree(d->d_send); cudaFree(d->d_recv); cudaFree(d->d_offsets); cudaFree(d->d_nacc); cudaFree(d->d_dense); cudaFree(d->d_tmp);
    unbind_current(d);
    if (d->peer_ptr) cudaIpcCloseMemHandle(d->peer_ptr);
    cudaFree(d->d_ipc);
    if (d->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(d->comm);
    delete d;
}

extern "C" int uavm_dist_rank(const uavm_dist* d) { return d ? d->rank : -1; }
extern "C" int uavm_dist_world(const uavm_dist* d) { return d ? d->world : 0; }

static int grow(uavm_ctx* ctx, void** p, size_t* cap, size_t need)
{
    if (*p && need <= *cap) return UAVM_OK;
    cudaFree(*p); *p = nullptr; *cap = 0;
    UAVM_CUDA(ctx, cudaMalloc(p, need));
    *cap = need;
    return UAVM_OK;
}

// Merge the RANSAC results of all ranks: pb holds this rank's shard of a global list of n_pairs_global pairs, pair p on
// rank p % world at local index p / world.  Every rank receives the full MatchPointPairs list (accepted pairs in global
// pair order, inliers in candidate order) — what one context running all pairs returns from uavm_pairbatch_collect.
// out == NULL: only counts.  world == 1 works without peers (and is how the single-GPU tests exercise this path).
extern "C" int uavm_pairbatch_allgather(uavm_ctx* ctx, uavm_dist* d, uavm_pairbatch* pb, int n_pairs_global, int min_inner_points,
                                        uavm_matchpointpairs* out, int cap, int* n_out, int* n_accepted_pairs)
{
    if (!ctx || !d || !n_out || n_pairs_global < 0) return UAVM_EINVAL;
    const int world = d->world, rank = d->rank;
    const int n_local = (n_pairs_global - rank + world - 1) / world;                 // pairs p = rank, rank + world, ...
    if ((pb ? pb->n_pairs : 0) != (n_local > 0 ? n_local : 0)) { UAVM_SET_ERR(ctx, "allgather: this rank must hold %d pairs (round-robin shard)", n_local); return UAVM_EINVAL; }
    if (pb && !pb->ransacked) { UAVM_SET_ERR(ctx, "allgather before ransac"); return UAVM_EINVAL; }
    UAVM_CUDA(ctx, cudaSetDevice(ctx->device));
    if (ctx->side_pending && !ctx->forked) { int rc = uavm_ctx_join(ctx); if (rc != UAVM_OK) return rc; }   // RANSAC may still run on the side stream
    const int n_slot = (n_pairs_global + world - 1) / world;                         // records per rank, padded
    if (n_slot == 0) { *n_out = 0; if (n_accepted_pairs) *n_accepted_pairs = 0; return UAVM_OK; }
    {
        int rc = grow(ctx, (void**)&d->d_send, &d->send_cap, (size_t)n_slot * sizeof(PairRecord)); if (rc != UAVM_OK) return rc;
        rc = grow(ctx, (void**)&d->d_recv, &d->recv_cap, (size_t)n_slot * world * sizeof(PairRecord)); if (rc != UAVM_OK) return rc;
        rc = grow(ctx, (void**)&d->d_offsets, &d->off_cap, ((size_t)n_pairs_global + 1) * sizeof(int32_t)); if (rc != UAVM_OK) return rc;
        if (!d->d_nacc) UAVM_CUDA(ctx, cudaMalloc(&d->d_nacc, sizeof(int32_t)));
    }
    k_dist_pack<<<n_slot, 256, 0, ctx->stream>>>(pb ? pb->d_pairs : nullptr, pb ? pb->n_pairs : 0, pb ? pb->d_cand_n : nullptr, pb ? pb->d_inlier : nullptr,
                                                 pb ? pb->d_cand_xy1 : nullptr, pb ? pb->d_cand_xy2 : nullptr, pb ? pb->d_cand_id1 : nullptr, pb ? pb->d_cand_id2 : nullptr,
                                                 pb ? pb->d_res : nullptr, d->d_send);
    UAVM_CHECK_LAUNCH(ctx);
    if (world == 1) UAVM_CUDA(ctx, cudaMemcpyAsync(d->d_recv, d->d_send, (size_t)n_slot * sizeof(PairRecord), cudaMemcpyDeviceToDevice, ctx->stream));
    else UAVM_NCCL(ctx, g_nccl.AllGather(d->d_send, d->d_recv, (size_t)n_slot * sizeof(PairRecord), ncclInt8, d->comm, ctx->stream));
    k_dist_offsets<<<1, 1024, 0, ctx->stream>>>(d->d_recv, n_pairs_global, world, n_slot, min_inner_points, d->d_offsets, d->d_nacc);
    UAVM_CHECK_LAUNCH(ctx);
    int32_t head[2] = {0, 0};
    UAVM_CUDA(ctx, cudaMemcpyAsync(&head[0], d->d_offsets + n_pairs_global, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    UAVM_CUDA(ctx, cudaMemcpyAsync(&head[1], d->d_nacc, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    UAVM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *n_out = head[0];
    if (n_accepted_pairs) *n_accepted_pairs = head[1];
    if (!out || head[0] == 0) return UAVM_OK;
    if (cap < head[0]) return UAVM_EINVAL;
    { int rc = grow(ctx, (void**)&d->d_dense, &d->dense_cap, (size_t)head[0] * sizeof(uavm_matchpointpairs)); if (rc != UAVM_OK) return rc; }
    k_dist_compact<<<n_pairs_global, 128, 0, ctx->stream>>>(d->d_recv, n_pairs_global, world, n_slot, min_inner_points, d->d_offsets, d->d_dense);
    UAVM_CHECK_LAUNCH(ctx);
    UAVM_CUDA(ctx, cudaMemcpyAsync(out, d->d_dense, (size_t)head[0] * sizeof(uavm_matchpointpairs), cudaMemcpyDeviceToHost, ctx->stream));
    UAVM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return UAVM_OK;
}

// Maps the root's mosaic buffer into every other rank's address space (CUDA IPC; cached per handle) and agrees on the outcome:
// *agreed = 1 when EVERY rank can write the root's buffer (all ranks must take the same path).  Collective; two small host
// round trips.
static int map_root_result(uavm_ctx* ctx, uavm_dist* d, uavm_canvas* cv, int root, int* agreed)
{
    *agreed = 0;
    if (!d->d_ipc) UAVM_CUDA(ctx, cudaMalloc(&d->d_ipc, 128));
    cudaIpcMemHandle_t h; memset(&h, 0, sizeof(h));
    int ok_local = 1;
    if (d->rank == root) {
        if (cudaIpcGetMemHandle(&h, cv->d_result) != cudaSuccess) { cudaGetLastError(); ok_local = 0; memset(&h, 0, sizeof(h)); }
        UAVM_CUDA(ctx, cudaMemcpyAsync(d->d_ipc, &h, sizeof(h), cudaMemcpyHostToDevice, ctx->stream));
    }
    UAVM_NCCL(ctx, g_nccl.Broadcast(d->d_ipc, d->d_ipc, sizeof(h), ncclInt8, root, d->comm, ctx->stream));
    if (d->rank != root) {
        UAVM_CUDA(ctx, cudaMemcpyAsync(&h, d->d_ipc, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
        UAVM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        bool zero = true;
        for (size_t i = 0; i < sizeof(h); i++) if (((const char*)&h)[i]) { zero = false; break; }
        if (zero || d->peer_failed) ok_local = 0;
        else if (!d->peer_ptr || memcmp(&h, &d->peer_handle, sizeof(h)) != 0) {
            if (d->peer_ptr) { unbind_current(d); cudaIpcCloseMemHandle(d->peer_ptr); d->peer_ptr = nullptr; }
            if (cudaIpcOpenMemHandle(&d->peer_ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); d->peer_ptr = nullptr; d->peer_failed = true; ok_local = 0; }
            else d->peer_handle = h;
        }
    }
    int* d_flag = reinterpret_cast<int*>(d->d_ipc + 64);
    const int cannot = ok_local ? 0 : 1;
    UAVM_CUDA(ctx, cudaMemcpyAsync(d_flag, &cannot, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    UAVM_NCCL(ctx, g_nccl.AllReduce(d_flag, d_flag, 1, ncclInt32, ncclSum, d->comm, ctx->stream));
    int n_cannot = 0;
    UAVM_CUDA(ctx, cudaMemcpyAsync(&n_cannot, d_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    UAVM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *agreed = n_cannot == 0;
    return UAVM_OK;
}

// Fused blend + gather: after this collective call the level-0 kernel of uavm_canvas_blend on every non-root rank stores its
// mosaic rectangle ALSO into the root's mosaic buffer over NVLink (peer memory mapped with CUDA IPC), tile by tile as the
// pixels are produced, so the transfer overlaps the blend and uavm_canvas_gather shrinks to its completion barrier.  Call it
// once per canvas, on every rank, after uavm_canvas_set_rect and before uavm_canvas_blend.  When the mapping is not possible
// (ranks sharing a process, IPC unavailable) nothing is bound and uavm_canvas_gather copies as before.  The binding refers to
// `d`: keep the uavm_dist alive as long as the canvas blends.
extern "C" int uavm_canvas_bind_root(uavm_ctx* ctx, uavm_dist* d, uavm_canvas* cv, int root)
{
    if (!ctx || !d || !cv || root < 0 || root >= d->world) return UAVM_EINVAL;
    if (cv->bound_dist) uavm_dist_forget_canvas(cv->bound_dist, cv);
    cv->bound_root = -1; cv->peer_result = nullptr; cv->bound_dist = nullptr;
    if (d->world == 1 || getenv("UAVM_GATHER_NCCL") || getenv("UAVM_GATHER_COPY")) return UAVM_OK;
    UAVM_CUDA(ctx, cudaSetDevice(ctx->device));
    { const int rc = uavm_canvas_ensure_result(ctx, cv); if (rc != UAVM_OK) return rc; }
    int agreed = 0;
    { const int rc = map_root_result(ctx, d, cv, root, &agreed); if (rc != UAVM_OK) return rc; }
    if (agreed) {
        if (d->bound_cv != cv) unbind_current(d);          // one bound canvas per uavm_dist: the mapping cache holds one buffer
        cv->bound_root = root; cv->bound_dist = d; d->bound_cv = cv;
        if (d->rank != root) cv->peer_result = static_cast<uint8_t*>(d->peer_ptr);
    }
    return UAVM_OK;
}

// Gather the finished rectangles of a sharded canvas on `root`: rects = world x 4 (x0, y0, x1, y1), rank r blended rects[r]
// (uavm_canvas_set_rect).  After the call the root's result holds the whole mosaic (uavm_canvas_get_result /
// uavm_canvas_copy_result_rows).  Full-width rectangles travel straight between the result buffers; 2-D pieces are packed
// on the sender and unpacked on the root.  Stream ordered on ctx's stream; the caller synchronises.
extern "C" int uavm_canvas_gather(uavm_ctx* ctx, uavm_dist* d, uavm_canvas* cv, const int32_t* rects, int root)
{
    if (!ctx || !d || !cv || !rects || root < 0 || root >= d->world) return UAVM_EINVAL;
    if (!cv->blended || !cv->d_result) { UAVM_SET_ERR(ctx, "canvas_gather before blend"); return UAVM_EINVAL; }
    UAVM_CUDA(ctx, cudaSetDevice(ctx->device));
    const int W = cv->result_w, H = cv->result_h;
    for (int r = 0; r < d->world; r++) {
        const int32_t* q = rects + 4 * r;
        if (q[0] < 0 || q[1] < 0 || q[2] > W || q[3] > H || q[0] > q[2] || q[1] > q[3]) return UAVM_EINVAL;
    }
    auto bytes = [&](int r) { const int32_t* q = rects + 4 * r; return (size_t)(q[2] - q[0]) * (size_t)(q[3] - q[1]) * 3; };
    auto full_width = [&](int r) { const int32_t* q = rects + 4 * r; return q[0] == 0 && q[2] == W; };
    if (d->world == 1) return UAVM_OK;
    int* d_flag = reinterpret_cast<int*>(d->d_ipc ? d->d_ipc + 64 : nullptr);
    // Fused path: the blend's level-0 kernel already wrote this rank's rectangle into the root's mosaic (uavm_canvas_bind_root);
    // what is left is the completion barrier, stream-ordered after the blend on every rank.
    if (cv->blend_bound_root == root && d_flag) {
        UAVM_CUDA(ctx, cudaMemsetAsync(d_flag, 0, sizeof(int), ctx->stream));
        UAVM_NCCL(ctx, g_nccl.AllReduce(d_flag, d_flag, 1, ncclInt32, ncclSum, d->comm, ctx->stream));
        return UAVM_OK;
    }
    // Direct path: every rank writes its rectangle straight into the root's result buffer over NVLink (the buffer is mapped
    // with CUDA IPC; one strided peer copy per rank, no packing, no staging on the root); a tiny all-reduce, stream-ordered
    // after the copies, tells the root that every rectangle has landed.  Falls back to grouped ncclSend / ncclRecv when the
    // mapping is not possible (all ranks in ONE process, or IPC unavailable).
    if (!getenv("UAVM_GATHER_NCCL")) {
        int agreed = 0;
        { const int rc = map_root_result(ctx, d, cv, root, &agreed); if (rc != UAVM_OK) return rc; }
        d_flag = reinterpret_cast<int*>(d->d_ipc + 64);
        if (agreed) {
            if (d->rank != root && bytes(d->rank) > 0) {
                const int32_t* q = rects + 4 * d->rank;
                const size_t off = ((size_t)q[1] * W + q[0]) * 3;
                UAVM_CUDA(ctx, cudaMemcpy2DAsync(static_cast<uint8_t*>(d->peer_ptr) + off, (size_t)W * 3, cv->d_result + off, (size_t)W * 3,
                                                 (size_t)(q[2] - q[0]) * 3, (size_t)(q[3] - q[1]), cudaMemcpyDeviceToDevice, ctx->stream));
            }
            UAVM_NCCL(ctx, g_nccl.AllReduce(d_flag, d_flag, 1, ncclInt32, ncclSum, d->comm, ctx->stream));     // completion barrier
            return UAVM_OK;
        }
    }
    if (d->rank != root) {
        const int32_t* q = rects + 4 * d->rank;
        if (bytes(d->rank) == 0) return UAVM_OK;
        const uint8_t* src = cv->d_result + ((size_t)q[1] * W + q[0]) * 3;
        if (!full_width(d->rank)) {
            int rc = grow(ctx, (void**)&d->d_tmp, &d->tmp_cap, bytes(d->rank)); if (rc != UAVM_OK) return rc;
            UAVM_CUDA(ctx, cudaMemcpy2DAsync(d->d_tmp, (size_t)(q[2] - q[0]) * 3, src, (size_t)W * 3, (size_t)(q[2] - q[0]) * 3, (size_t)(q[3] - q[1]),
                                             cudaMemcpyDeviceToDevice, ctx->stream));
            src = d->d_tmp;
        }
        UAVM_NCCL(ctx, g_nccl.Send(src, bytes(d->rank), ncclInt8, root, d->comm, ctx->stream));
        return UAVM_OK;
    }
    size_t tmp_need = 0;
    for (int r = 0; r < d->world; r++) if (r != root && !full_width(r)) tmp_need += (bytes(r) + 255) & ~(size_t)255;
    if (tmp_need) { int rc = grow(ctx, (void**)&d->d_tmp, &d->tmp_cap, tmp_need); if (rc != UAVM_OK) return rc; }
    UAVM_NCCL(ctx, g_nccl.GroupStart());
    size_t off = 0;
    for (int r = 0; r < d->world; r++) {
        if (r == root || bytes(r) == 0) continue;
        const int32_t* q = rects + 4 * r;
        uint8_t* dst = full_width(r) ? cv->d_result + (size_t)q[1] * W * 3 : d->d_tmp + off;
        if (!full_width(r)) off += (bytes(r) + 255) & ~(size_t)255;
        ncclResult_t rr = g_nccl.Recv(dst, bytes(r), ncclInt8, r, d->comm, ctx->stream);
        if (rr != ncclSuccess) { g_nccl.GroupEnd(); UAVM_SET_ERR(ctx, "ncclRecv -> %s", g_nccl.GetErrorString(rr)); return UAVM_EFAIL; }
    }
    UAVM_NCCL(ctx, g_nccl.GroupEnd());
    off = 0;
    for (int r = 0; r < d->world; r++) {
        if (r == root || bytes(r) == 0 || full_width(r)) continue;
        const int32_t* q = rects + 4 * r;
        UAVM_CUDA(ctx, cudaMemcpy2DAsync(cv->d_result + ((size_t)q[1] * W + q[0]) * 3, (size_t)W * 3, d->d_tmp + off, (size_t)(q[2] - q[0]) * 3,
                                         (size_t)(q[2] - q[0]) * 3, (size_t)(q[3] - q[1]), cudaMemcpyDeviceToDevice, ctx->stream));
        off += (bytes(r) + 255) & ~(size_t)255;
    }
    return UAVM_OK;
}

// root of the canvas' binding (uavm_canvas_bind_root), -1 when nothing is bound
extern "C" int uavm_canvas_bound_root(const uavm_canvas* cv) { return cv ? cv->bound_root : -1; }

// replicate `bytes` of device memory from `root` to every rank (descriptor pools, transforms): ncclBroadcast on ctx's stream
extern "C" int uavm_dist_broadcast(uavm_ctx* ctx, uavm_dist* d, void* device_buf, int64_t bytes, int root)
{
    if (!ctx || !d || !device_buf || bytes < 0 || root < 0 || root >= d->world) return UAVM_EINVAL;
    if (bytes == 0 || d->world == 1) return UAVM_OK;
    UAVM_CUDA(ctx, cudaSetDevice(ctx->device));
    UAVM_NCCL(ctx, g_nccl.Broadcast(device_buf, device_buf, (size_t)bytes, ncclInt8, root, d->comm, ctx->stream));
    return UAVM_OK;
}
