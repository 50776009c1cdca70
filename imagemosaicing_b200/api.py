"""Host-side mirror of the reference's entry points for the match -> RANSAC -> warp/blend path,
implemented on the C ABI of libuavmosaic.so (include/uavm.h).  Names follow the reference:

    match()                <- FlannBasedMatcher().match            M/MosaicWithoutPos.cpp:5108-5110
    select_match_pairs()   <- std::sort + SelectMatchPairs          :5111, :5146-5153, :4977-5028
    ransac2d()             <- Ransac2D                              M/mosaicimage.h:1729-2035
    PairBatch              <- the per-pair body of GetMatchedPairsOneToAllSIFTThread (:5083-5227), batched
    align_affine()         <- BundleAdjustmentSparse                :6971-7202
    Canvas                 <- LaplacianPyramidBlending              M/MosaicImage.cpp:2205-2510

numpy arrays are host buffers (copied by the library); torch CUDA tensors are passed as device
pointers.  There is no CPU fallback anywhere in this module.
"""
import ctypes as C
import numpy as np

from . import _lib as L
from ._lib import (SfPoint, DMatch, RansacResult, MatchPointPairs, ImageTransform, ChipLayout, CanvasLayout,
                   f32p, i32p, u32p, u8p)


class UavmError(RuntimeError):
    pass


def _is_torch_cuda(x):
    return hasattr(x, "is_cuda") and x.is_cuda


def _host(x):
    """torch CPU tensors (e.g. pinned host buffers) are viewed as numpy arrays without a copy."""
    if hasattr(x, "is_cuda") and not x.is_cuda:
        return x.numpy()
    return x


def _ptr(x, ctype):
    if _is_torch_cuda(x):
        return C.cast(C.c_void_p(x.data_ptr()), ctype)
    return x.ctypes.data_as(ctype)


class Context:
    """One per process/GPU (uavm_ctx)."""

    def __init__(self, device=0, stream=None):
        self._h = C.c_void_p()
        self.device = int(device)
        rc = L.lib().uavm_ctx_create(int(device), C.byref(self._h))
        if rc != 0:
            raise UavmError(f"uavm_ctx_create({device}) failed with {rc}: no sm_100 CUDA device (no CPU fallback)")
        if stream is not None:
            self.set_stream(stream)

    def set_stream(self, stream):
        ptr = stream.cuda_stream if hasattr(stream, "cuda_stream") else int(stream)
        self.check(L.lib().uavm_ctx_set_stream(self._h, C.c_void_p(ptr)))

    def check(self, rc):
        if rc != 0:
            raise UavmError(f"uavm error {rc}: {L.lib().uavm_last_error(self._h).decode()}")

    def sync(self):
        self.check(L.lib().uavm_ctx_sync(self._h))

    def fork(self):
        self.check(L.lib().uavm_ctx_fork(self._h))

    def unfork(self):
        self.check(L.lib().uavm_ctx_unfork(self._h))

    def join(self):
        self.check(L.lib().uavm_ctx_join(self._h))

    @property
    def launch_count(self):
        return int(L.lib().uavm_ctx_launch_count(self._h))

    @property
    def sm_count(self):
        return int(L.lib().uavm_ctx_sm_count(self._h))

    def close(self):
        if self._h:
            L.lib().uavm_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class FeatureSet:
    """Device-resident descriptors + keypoints of a set of images (uavm_featureset)."""

    def __init__(self, ctx, n_keypoints):
        self.ctx = ctx
        self.n = np.ascontiguousarray(n_keypoints, np.int32)
        self._h = C.c_void_p()
        ctx.check(L.lib().uavm_featureset_create(ctx._h, len(self.n), _ptr(self.n, i32p), C.byref(self._h)))

    def upload(self, image, desc, kp_xy=None):
        """desc: (n,128) uint8 or float32 (integer valued) numpy array or torch CUDA tensor; kp_xy: (n,2) f32."""
        desc = _host(desc); kp_xy = _host(kp_xy) if kp_xy is not None else None
        dev = 1 if _is_torch_cuda(desc) else 0
        if dev:
            is_u8 = str(desc.dtype) == "torch.uint8"
            assert desc.is_contiguous()
            if kp_xy is not None:
                assert _is_torch_cuda(kp_xy) and kp_xy.is_contiguous()
        else:
            is_u8 = desc.dtype == np.uint8
            desc = np.ascontiguousarray(desc, np.uint8 if is_u8 else np.float32)
            if kp_xy is not None:
                kp_xy = np.ascontiguousarray(kp_xy, np.float32)
        assert desc.shape[0] == self.n[image] and desc.shape[1] == 128
        kp_ptr = _ptr(kp_xy, f32p) if kp_xy is not None else None
        if is_u8:
            rc = L.lib().uavm_featureset_upload_u8(self.ctx._h, self._h, int(image), _ptr(desc, u8p), kp_ptr, dev)
        else:
            rc = L.lib().uavm_featureset_upload_f32(self.ctx._h, self._h, int(image), _ptr(desc, f32p), kp_ptr, dev)
        self.ctx.check(rc)
        # host sources are copied asynchronously (pinned memory: the DMA runs after this call returns): keep one reference PER
        # IMAGE until that image is uploaded again or the object is closed, so a caller may pass temporaries
        if not hasattr(self, "_keep"):
            self._keep = {}
        self._keep[int(image)] = (desc, kp_xy)

    def close(self):
        if self._h:
            L.lib().uavm_featureset_destroy(self.ctx._h, self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Sift:
    """SIFT detect + compute on the GPU (uavm_sift; replaces SiftExtraction_Thread's SIFT(2000, 3, 0.01, 20),
    M/MosaicWithoutPos.cpp:4852-4872).  One object per image size."""

    def __init__(self, ctx, img_w, img_h, nfeatures=2000, n_octave_layers=3, contrast_threshold=0.01, edge_threshold=20.0, sigma=1.6):
        self.ctx = ctx; self.w = int(img_w); self.h = int(img_h)
        self._h = C.c_void_p()
        ctx.check(L.lib().uavm_sift_create(ctx._h, self.w, self.h, int(nfeatures), int(n_octave_layers), C.c_double(contrast_threshold),
                                           C.c_double(edge_threshold), C.c_double(sigma), C.byref(self._h)))

    def detect_and_compute(self, bgr, cap=1 << 17):
        """bgr: (h, w, 3) uint8 numpy array or torch CUDA tensor -> (keypoints structured array, descriptors (n, 128) f32)."""
        bgr = _host(bgr)
        dev = 1 if _is_torch_cuda(bgr) else 0
        if dev:
            assert bgr.is_contiguous(); step = bgr.stride(0)
        else:
            bgr = np.ascontiguousarray(bgr, np.uint8); step = bgr.strides[0]
        assert bgr.shape[0] == self.h and bgr.shape[1] == self.w and bgr.shape[2] == 3
        kp = (L.KeyPoint * cap)(); desc = np.zeros((cap, 128), np.float32); n = C.c_int(0)
        self.ctx.check(L.lib().uavm_sift_detect_and_compute(self.ctx._h, self._h, _ptr(bgr, u8p), int(step), dev, kp, _ptr(desc, f32p), cap, C.byref(n)))
        dt = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])
        return np.frombuffer(kp, dtype=dt)[:n.value].copy(), desc[:n.value].copy()

    def close(self):
        if self._h:
            L.lib().uavm_sift_destroy(self.ctx._h, self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Jpeg:
    """nvJPEG decoder handle (uavm_jpeg): JPEG bytes in, BGR pixels in HBM out."""

    def __init__(self, ctx, backend=0):
        self.ctx = ctx
        self._h = C.c_void_p()
        ctx.check(L.lib().uavm_jpeg_create(ctx._h, int(backend), C.byref(self._h)))

    def decode(self, data, out):
        """data: bytes / uint8 numpy array with the JPEG stream; out: torch CUDA uint8 tensor (h, w, 3), contiguous."""
        buf = np.frombuffer(data, np.uint8) if isinstance(data, (bytes, bytearray)) else np.ascontiguousarray(data, np.uint8)
        assert _is_torch_cuda(out) and out.is_contiguous()
        self.ctx.check(L.lib().uavm_jpeg_decode_bgr(self.ctx._h, self._h, _ptr(buf, u8p), C.c_int64(buf.size), C.c_void_p(out.data_ptr()),
                                                    int(out.stride(0)), int(out.shape[1]), int(out.shape[0])))

    def set_canvas_image(self, cv, image, data):
        buf = np.frombuffer(data, np.uint8) if isinstance(data, (bytes, bytearray)) else np.ascontiguousarray(data, np.uint8)
        self.ctx.check(L.lib().uavm_canvas_set_image_jpeg(self.ctx._h, cv._h, self._h, int(image), _ptr(buf, u8p), C.c_int64(buf.size)))

    def set_canvas_images(self, cv, first, datas):
        """uavm_canvas_set_images_jpeg: frames [first, first + len(datas)) decoded as ONE nvJPEG batch."""
        bufs = [np.frombuffer(d, np.uint8) if isinstance(d, (bytes, bytearray)) else np.ascontiguousarray(d, np.uint8).reshape(-1) for d in datas]
        ptrs = (C.c_void_p * len(bufs))(*[b.ctypes.data for b in bufs])
        sizes = (C.c_int64 * len(bufs))(*[b.size for b in bufs])
        self.ctx.check(L.lib().uavm_canvas_set_images_jpeg(self.ctx._h, cv._h, self._h, int(first), len(bufs), ptrs, sizes))
        self._keep = bufs                      # the decoder reads the streams asynchronously

    def set_threads(self, n):
        """host threads of set_canvas_images: n >= 1, 0 = default, -1 = nvJPEG's own batched decoder."""
        self.ctx.check(L.lib().uavm_jpeg_set_threads(self._h, int(n)))

    @property
    def hw_engines(self):
        return int(L.lib().uavm_jpeg_hw_engines(self._h))

    def close(self):
        if self._h:
            L.lib().uavm_jpeg_destroy(self.ctx._h, self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PairBatch:
    """match -> select -> RANSAC for a list of (query image, train image) pairs (uavm_pairbatch)."""

    def __init__(self, ctx, fs, pairs):
        self.ctx = ctx
        self.fs = fs
        self.pairs = np.ascontiguousarray(pairs, np.int32).reshape(-1, 2)
        self._h = C.c_void_p()
        ctx.check(L.lib().uavm_pairbatch_create(ctx._h, fs._h, len(self.pairs), _ptr(self.pairs, i32p), C.byref(self._h)))

    def match(self):
        self.ctx.check(L.lib().uavm_pairbatch_match(self.ctx._h, self._h))

    def select(self, width, height, grid_x=3, grid_y=3, max_num=400, frac=0.3):
        self.ctx.check(L.lib().uavm_pairbatch_select(self.ctx._h, self._h, int(width), int(height), grid_x, grid_y,
                                                     max_num, C.c_double(frac)))

    def ransac(self, ransac_dist=2.5, sample_times=1000, seeds=None, base_seed=0):
        sp = None
        if seeds is not None:
            seeds = np.ascontiguousarray(seeds, np.uint32)
            assert len(seeds) == len(self.pairs)
            sp = _ptr(seeds, u32p)
        self.ctx.check(L.lib().uavm_pairbatch_ransac(self.ctx._h, self._h, C.c_float(ransac_dist), int(sample_times),
                                                     sp, C.c_uint32(base_seed)))

    def matches(self, pair):
        nq = int(self.fs.n[self.pairs[pair, 0]])
        out = (DMatch * max(nq, 1))()
        n = C.c_int(0)
        self.ctx.check(L.lib().uavm_pairbatch_get_matches(self.ctx._h, self._h, int(pair), out, nq, C.byref(n)))
        a = np.frombuffer(out, dtype=np.dtype([("queryIdx", "<i4"), ("trainIdx", "<i4"), ("imgIdx", "<i4"), ("distance", "<f4")]))
        return a[:n.value].copy()

    def candidates(self, pair, cap=1024):
        p1 = (SfPoint * cap)(); p2 = (SfPoint * cap)(); n = C.c_int(0)
        self.ctx.check(L.lib().uavm_pairbatch_get_candidates(self.ctx._h, self._h, int(pair), p1, p2, cap, C.byref(n)))
        dt = np.dtype([("x", "<f4"), ("y", "<f4"), ("id", "<i4")])
        return np.frombuffer(p1, dtype=dt)[:n.value].copy(), np.frombuffer(p2, dtype=dt)[:n.value].copy()

    def ransac_result(self, pair, cap=1024):
        mask = np.zeros(cap, np.uint8); res = RansacResult()
        self.ctx.check(L.lib().uavm_pairbatch_get_ransac(self.ctx._h, self._h, int(pair), _ptr(mask, u8p), cap, C.byref(res)))
        return mask, res

    def collect(self, min_inner_points=30, cap=None):
        n = C.c_int(0); acc = C.c_int(0)
        self.ctx.check(L.lib().uavm_pairbatch_collect(self.ctx._h, self._h, int(min_inner_points), None, 0, C.byref(n), C.byref(acc)))
        out = (MatchPointPairs * max(n.value, 1))()
        self.ctx.check(L.lib().uavm_pairbatch_collect(self.ctx._h, self._h, int(min_inner_points), out, n.value, C.byref(n), C.byref(acc)))
        return out, n.value, acc.value

    def close(self):
        if self._h:
            L.lib().uavm_pairbatch_destroy(self.ctx._h, self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Dist:
    """NCCL communicator of this rank behind the C ABI (uavm_dist; csrc/dist.cu).  `Dist.from_torch(ctx)` bootstraps it in a
    torchrun job (rank 0 creates the NCCL id, torch.distributed carries it to the other ranks); a C++ host hands the id
    around by its own means.  world == 1 needs no peers."""

    ID_BYTES = 128

    @staticmethod
    def _one_nccl_per_process():
        """The library picks up an NCCL that is already loaded; a Python process that will also import torch must load torch's
        bundled NCCL first (torch's libraries are linked against that version)."""
        try:
            import torch  # noqa: F401
        except Exception:
            pass

    def __init__(self, ctx, rank, world, unique_id):
        self._one_nccl_per_process()
        self.ctx = ctx; self.rank = int(rank); self.world = int(world)
        idb = (C.c_uint8 * self.ID_BYTES).from_buffer_copy(bytes(unique_id))
        self._h = C.c_void_p()
        ctx.check(L.lib().uavm_dist_init(ctx._h, self.rank, self.world, idb, self.ID_BYTES, C.byref(self._h)))

    @staticmethod
    def unique_id():
        Dist._one_nccl_per_process()
        idb = (C.c_uint8 * Dist.ID_BYTES)()
        rc = L.lib().uavm_dist_unique_id(idb, Dist.ID_BYTES)
        if rc != 0:
            raise UavmError(f"uavm_dist_unique_id failed with {rc} (libnccl.so.2 missing?)")
        return bytes(idb)

    @staticmethod
    def from_torch(ctx):
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return Dist(ctx, 0, 1, bytes(Dist.ID_BYTES))            # a single rank needs no communicator
        rank, world = dist.get_rank(), dist.get_world_size()
        box = [Dist.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        return Dist(ctx, rank, world, box[0])

    def allgather_matches(self, pb, n_pairs_global, min_inner_points=30):
        """-> (MatchPointPairs array of ALL pairs in global pair order, n, n_accepted_pairs); identical on every rank."""
        h = pb._h if pb is not None else None
        n = C.c_int(0); acc = C.c_int(0)
        self.ctx.check(L.lib().uavm_pairbatch_allgather(self.ctx._h, self._h, h, int(n_pairs_global), int(min_inner_points), None, 0, C.byref(n), C.byref(acc)))
        out = (MatchPointPairs * max(n.value, 1))()
        self.ctx.check(L.lib().uavm_pairbatch_allgather(self.ctx._h, self._h, h, int(n_pairs_global), int(min_inner_points), out, n.value, C.byref(n), C.byref(acc)))
        return out, n.value, acc.value

    def bind_canvas_root(self, cv, root=0):
        """uavm_canvas_bind_root (collective, before blend): the blend's level-0 kernel of every other rank also writes its rectangle
        into root's mosaic over NVLink; gather_canvas is then only the completion barrier."""
        self.ctx.check(L.lib().uavm_canvas_bind_root(self.ctx._h, self._h, cv._h, int(root)))
        return int(L.lib().uavm_canvas_bound_root(cv._h)) == int(root)

    def gather_canvas(self, cv, rects, root=0):
        """rects: (world, 4) int32 (x0, y0, x1, y1) per rank; afterwards root's canvas result is the whole mosaic."""
        r = np.ascontiguousarray(rects, np.int32).reshape(-1, 4)
        assert len(r) == self.world
        self.ctx.check(L.lib().uavm_canvas_gather(self.ctx._h, self._h, cv._h, _ptr(r, i32p), int(root)))

    def broadcast(self, tensor, root=0):
        """torch CUDA tensor, replicated from root in place (ncclBroadcast on the ctx stream)."""
        assert _is_torch_cuda(tensor) and tensor.is_contiguous()
        self.ctx.check(L.lib().uavm_dist_broadcast(self.ctx._h, self._h, C.c_void_p(tensor.data_ptr()), C.c_int64(tensor.numel() * tensor.element_size()), int(root)))

    def close(self):
        if self._h:
            L.lib().uavm_dist_destroy(self.ctx._h, self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- single-pair seams (host buffers in, host buffers out) -------------------------------------
def match(ctx, desc1, desc2):
    """Exact L2 1-NN of every row of desc1 in desc2 -> structured array (queryIdx, trainIdx, imgIdx, distance)."""
    d1 = np.ascontiguousarray(desc1, np.float32); d2 = np.ascontiguousarray(desc2, np.float32)
    out = (DMatch * max(len(d1), 1))()
    ctx.check(L.lib().uavm_match(ctx._h, _ptr(d1, f32p), len(d1), _ptr(d2, f32p), len(d2), out))
    a = np.frombuffer(out, dtype=np.dtype([("queryIdx", "<i4"), ("trainIdx", "<i4"), ("imgIdx", "<i4"), ("distance", "<f4")]))
    return a[:len(d1)].copy()


def select_match_pairs(ctx, matches, kp1_xy, kp2_xy, width, height, grid_x=3, grid_y=3, max_num=400, frac=0.3):
    m = np.ascontiguousarray(matches)
    kp1 = np.ascontiguousarray(kp1_xy, np.float32); kp2 = np.ascontiguousarray(kp2_xy, np.float32)
    cap = 1024
    p1 = (SfPoint * cap)(); p2 = (SfPoint * cap)(); n = C.c_int(0)
    ctx.check(L.lib().uavm_select(ctx._h, C.cast(m.ctypes.data, C.POINTER(DMatch)), len(m), _ptr(kp1, f32p), len(kp1),
                                  _ptr(kp2, f32p), len(kp2), int(width), int(height), grid_x, grid_y, max_num,
                                  C.c_double(frac), p1, p2, cap, C.byref(n)))
    dt = np.dtype([("x", "<f4"), ("y", "<f4"), ("id", "<i4")])
    return np.frombuffer(p1, dtype=dt)[:n.value].copy(), np.frombuffer(p2, dtype=dt)[:n.value].copy()


def ransac2d(ctx, xy1, xy2, ransac_dist=2.5, sample_times=1000, seed=1):
    """Ransac2D on candidate pairs (n,2)+(n,2).  Returns (ok, inlier_mask, H[9], RansacResult)."""
    xy1 = np.ascontiguousarray(xy1, np.float32); xy2 = np.ascontiguousarray(xy2, np.float32)
    n = len(xy1)
    a = (SfPoint * max(n, 1))(); b = (SfPoint * max(n, 1))()
    for i in range(n):
        a[i].x, a[i].y, a[i].id = float(xy1[i, 0]), float(xy1[i, 1]), i
        b[i].x, b[i].y, b[i].id = float(xy2[i, 0]), float(xy2[i, 1]), i
    in1 = (SfPoint * max(n, 1))(); in2 = (SfPoint * max(n, 1))(); res = RansacResult()
    ctx.check(L.lib().uavm_ransac2d(ctx._h, a, b, n, C.c_float(ransac_dist), int(sample_times), C.c_uint32(seed),
                                    in1, in2, n, C.byref(res)))
    mask = np.zeros(n, np.uint8)
    for k in range(res.n_inliers):
        mask[in1[k].id] = 1
    return res.ok, mask, np.array(list(res.H), np.float32), res


def align_affine_rot(matches, n_matches, n_images, fixed, weight=100.0, iterations=10):
    """f3: SparseAffineRotConstraint (M/MosaicWithoutPos.cpp:6302-6808) on a MatchPointPairs ctypes array whose fixed flags are
    set; `fixed`: (n_images,) 0/1.  Returns (n_images, 9) float32 transforms."""
    init = (ImageTransform * n_images)(); out = (ImageTransform * n_images)()
    for i in range(n_images):
        for t in range(9):
            init[i].h.m[t] = 1.0 if t in (0, 4, 8) else 0.0
        init[i].fixed = int(fixed[i])
    rc = L.lib().uavm_align_affine_rot(matches, int(n_matches), init, int(n_images), int(sum(int(f) for f in fixed)), C.c_float(weight), int(iterations), out)
    if rc != 0:
        raise UavmError(f"uavm_align_affine_rot failed with {rc}")
    return np.array([[out[i].h.m[t] for t in range(9)] for i in range(n_images)], np.float32)


# ---- warp / seam masks / blend --------------------------------------------------------------------
def canvas_layout(H, keep, img_w, img_h):
    """Canvas sizing + chip boxes (M/MosaicImage.cpp:2233-2348), host side."""
    H = np.ascontiguousarray(H, np.float32).reshape(-1, 9)
    n = len(H)
    kp = None
    if keep is not None:
        keep = np.ascontiguousarray(keep, np.int32); kp = _ptr(keep, i32p)
    canvas = CanvasLayout(); chips = (ChipLayout * n)()
    rc = L.lib().uavm_canvas_layout_compute(_ptr(H, f32p), kp, n, int(img_w), int(img_h), C.byref(canvas), chips)
    if rc != 0:
        raise UavmError(f"uavm_canvas_layout_compute failed: {rc}")
    return canvas, chips


class Canvas:
    """Warp + seam masks + blend of N frames with transforms H (uavm_canvas; LaplacianPyramidBlending)."""

    def __init__(self, ctx, H, img_w, img_h, keep=None):
        self.ctx = ctx
        self.H = np.ascontiguousarray(H, np.float32).reshape(-1, 9)
        self.n = len(self.H); self.img_w = int(img_w); self.img_h = int(img_h)
        kp = None
        if keep is not None:
            keep = np.ascontiguousarray(keep, np.int32); kp = _ptr(keep, i32p)
        self._h = C.c_void_p()
        ctx.check(L.lib().uavm_canvas_create(ctx._h, self.n, self.img_w, self.img_h, _ptr(self.H, f32p), kp, C.byref(self._h)))
        self.layout = CanvasLayout(); self.chips = (ChipLayout * self.n)()
        ctx.check(L.lib().uavm_canvas_get_layout(self._h, C.byref(self.layout), self.chips))

    @property
    def source_layout(self):
        """3: frames stay BGR in HBM (no conversion pass), 4: BGRA pool."""
        return int(L.lib().uavm_canvas_source_layout(self._h))

    def source_frame(self, image):
        """uavm_canvas_image_ptr: zero-copy torch view (h, w, 3) of source frame `image` in a BGR pool."""
        import torch
        p = C.c_void_p(); step = C.c_int()
        self.ctx.check(L.lib().uavm_canvas_image_ptr(self._h, int(image), C.byref(p), C.byref(step)))
        class _View:
            pass
        v = _View()
        v.__cuda_array_interface__ = {"shape": (self.img_h, self.img_w, 3), "typestr": "|u1", "data": (int(p.value), False), "version": 2,
                                      "strides": (int(step.value), 3, 1)}
        return torch.as_tensor(v, device=f"cuda:{self.ctx.device}")

    def set_image(self, image, bgr):
        """bgr: (h, w, 3) uint8 numpy array (host) or torch CUDA tensor (device)."""
        bgr = _host(bgr)
        dev = 1 if _is_torch_cuda(bgr) else 0
        if dev:
            assert bgr.is_contiguous(); step = bgr.stride(0)
        else:
            bgr = np.ascontiguousarray(bgr, np.uint8); step = bgr.strides[0]
        assert bgr.shape[0] == self.img_h and bgr.shape[1] == self.img_w and bgr.shape[2] == 3
        self.ctx.check(L.lib().uavm_canvas_set_image(self.ctx._h, self._h, int(image), _ptr(bgr, u8p), int(step), dev))
        if not hasattr(self, "_keep"):
            self._keep = {}
        self._keep[int(image)] = bgr       # per image: the copy is asynchronous for pinned host frames (see FeatureSet.upload)

    def set_band(self, y0, y1, halo=0):
        """Multi-GPU canvas sharding: produce only canvas rows [y0, y1); see uavm_canvas_set_band (`halo` is ignored)."""
        self.ctx.check(L.lib().uavm_canvas_set_band(self.ctx._h, self._h, int(y0), int(y1), int(halo)))

    def set_rect(self, x0, y0, x1, y1):
        """Multi-GPU canvas sharding: produce only the canvas rectangle [x0, x1) x [y0, y1) (even edges); uavm_canvas_set_rect."""
        self.ctx.check(L.lib().uavm_canvas_set_rect(self.ctx._h, self._h, int(x0), int(y0), int(x1), int(y1)))

    def is_active(self, image):
        return bool(L.lib().uavm_canvas_is_active(self._h, int(image)))

    def warp(self, first=None, count=None):
        """K5 for every kept frame, or for frames [first, first + count) (streaming callers warp each group of
        frames as soon as it has been set, while the next frames are still crossing PCIe)."""
        if first is None:
            self.ctx.check(L.lib().uavm_canvas_warp(self.ctx._h, self._h))
        else:
            self.ctx.check(L.lib().uavm_canvas_warp_range(self.ctx._h, self._h, int(first), int(count)))

    def warp_for_blend(self):
        """K5 restricted to what the blend reads (needs seam_masks() first); uavm_canvas_warp_for_blend."""
        self.ctx.check(L.lib().uavm_canvas_warp_for_blend(self.ctx._h, self._h))

    def seam_masks(self):
        self.ctx.check(L.lib().uavm_canvas_seam_masks(self.ctx._h, self._h))

    def blend(self, num_bands=5):
        self.ctx.check(L.lib().uavm_canvas_blend(self.ctx._h, self._h, int(num_bands)))

    def paste(self):
        self.ctx.check(L.lib().uavm_canvas_paste(self.ctx._h, self._h))

    def chip(self, image):
        c = self.chips[image]
        px = np.zeros((c.chip_h, c.chip_w, 3), np.uint8); mask = np.zeros((c.chip_h, c.chip_w), np.uint8)
        self.ctx.check(L.lib().uavm_canvas_get_chip(self.ctx._h, self._h, int(image), _ptr(px, u8p), px.strides[0],
                                                    _ptr(mask, u8p), mask.strides[0]))
        return px, mask

    def result(self):
        rw = C.c_int(0); rh = C.c_int(0)
        self.ctx.check(L.lib().uavm_canvas_result_size(self._h, C.byref(rw), C.byref(rh)))
        out = np.zeros((rh.value, rw.value, 3), np.uint8)
        mask = np.zeros((rh.value, rw.value), np.uint8)
        self.ctx.check(L.lib().uavm_canvas_get_result(self.ctx._h, self._h, _ptr(out, u8p), out.strides[0],
                                                      _ptr(mask, u8p), mask.strides[0]))
        return out, mask

    def copy_result_rows(self, y0, y1, dst):
        """Rows [y0, y1) of the result into `dst` (torch CUDA uint8 tensor or numpy array, (y1-y0, W, 3))."""
        dev = 1 if _is_torch_cuda(dst) else 0
        self.ctx.check(L.lib().uavm_canvas_copy_result_rows(self.ctx._h, self._h, int(y0), int(y1), _ptr(dst, u8p), dev))

    def copy_result_rect(self, x0, y0, x1, y1, dst):
        """Rectangle [x0, x1) x [y0, y1) of the result, dense, into `dst` (torch CUDA uint8 tensor or numpy array)."""
        dev = 1 if _is_torch_cuda(dst) else 0
        self.ctx.check(L.lib().uavm_canvas_copy_result_rect(self.ctx._h, self._h, int(x0), int(y0), int(x1), int(y1), _ptr(dst, u8p), dev))

    def close(self):
        if self._h:
            L.lib().uavm_canvas_destroy(self.ctx._h, self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _image_array(images):
    n = len(images)
    ims = [np.ascontiguousarray(i, np.uint8) for i in images]
    arr = (L.Image * n)()
    for i, im in enumerate(ims):
        arr[i].width, arr[i].height, arr[i].nChannels, arr[i].widthStep = im.shape[1], im.shape[0], 3, im.strides[0]
        arr[i].imageData = im.ctypes.data
    return ims, arr


def _param(param):
    P = L.Param()
    L.lib().uavm_param_default(C.byref(P))
    for k, v in (param or {}).items():
        setattr(P, k, v)
    return P


def _take_result(res, tr, n):
    buf = (C.c_uint8 * (res.widthStep * res.height)).from_address(res.imageData)
    out = np.frombuffer(buf, np.uint8).reshape(res.height, res.widthStep)[:, :res.width * 3].reshape(res.height, res.width, 3).copy()
    L.lib().uavm_free(C.c_void_p(res.imageData))
    T = np.array([[tr[i].h.m[t] for t in range(9)] for i in range(n)], np.float32)
    fixed = np.array([tr[i].fixed for i in range(n)], np.int32)
    return out, T, fixed


def mosaic_images(ctx, images, descs, kps, param=None, scale=1.0, return_matches=False):
    """MosaicVavImages-shaped entry (uavm_mosaic_images): images = list of (h,w,3) u8 BGR frames, descs = list of
    (n_i,128) f32/u8 descriptors, kps = list of (n_i,2) f32.  Returns (mosaic u8 (H,W,3), transforms (n,9) f32, fixed (n,))
    and, with return_matches, the accepted inlier matches (MatchPointPairs array: what the reference writes to
    feature_temp/matchPairs.match)."""
    n = len(images)
    ims, arr = _image_array(images)
    ds = [np.ascontiguousarray(d, np.float32) for d in descs]
    ks = [np.ascontiguousarray(k, np.float32) for k in kps]
    dp = (f32p * n)(*[_ptr(d, f32p) for d in ds]); kp = (f32p * n)(*[_ptr(k, f32p) for k in ks])
    nk = np.array([len(d) for d in ds], np.int32)
    P = _param(param)
    res = L.Image(); nm = C.c_int(0); tr = (ImageTransform * n)()
    pp = C.POINTER(MatchPointPairs)(); npairs = C.c_int(0)
    rc = L.lib().uavm_mosaic_images_ex(ctx._h, arr, n, dp, kp, _ptr(nk, i32p), C.byref(P), C.c_float(scale), C.byref(res), C.byref(nm), tr,
                                       C.byref(pp), C.byref(npairs))
    matches = None
    if npairs.value > 0:
        matches = (MatchPointPairs * npairs.value)()
        C.memmove(matches, pp, C.sizeof(MatchPointPairs) * npairs.value)
        L.lib().uavm_free(C.cast(pp, C.c_void_p))
    ctx.check(rc)
    out = _take_result(res, tr, n)
    return out + (matches,) if return_matches else out


def mosaic_images_sift(ctx, images, param=None, scale=1.0):
    """MosaicVavImages' own shape (uavm_mosaic_images_sift): images in, (mosaic, transforms, fixed, matches) out; SIFT runs on
    the GPU with the reference's parameters."""
    n = len(images)
    ims, arr = _image_array(images)
    P = _param(param)
    res = L.Image(); nm = C.c_int(0); tr = (ImageTransform * n)()
    pp = C.POINTER(MatchPointPairs)(); npairs = C.c_int(0)
    rc = L.lib().uavm_mosaic_images_sift(ctx._h, arr, n, C.byref(P), C.c_float(scale), C.byref(res), C.byref(nm), tr, C.byref(pp), C.byref(npairs))
    matches = None
    if npairs.value > 0:
        matches = (MatchPointPairs * npairs.value)()
        C.memmove(matches, pp, C.sizeof(MatchPointPairs) * npairs.value)
        L.lib().uavm_free(C.cast(pp, C.c_void_p))
    ctx.check(rc)
    return _take_result(res, tr, n) + (matches,)


def mosaic_from_matches(ctx, images, matches, param=None, scale=1.0):
    """The loadMatchPairs = 1 path (M/MosaicWithoutPos.cpp:4465-4477) without the stdin prompt: `matches` (a ctypes array of
    MatchPointPairs, e.g. from read_match_file) replaces feature matching."""
    n = len(images)
    ims, arr = _image_array(images)
    P = _param(param)
    res = L.Image(); nm = C.c_int(0); tr = (ImageTransform * n)()
    ctx.check(L.lib().uavm_mosaic_from_matches(ctx._h, arr, n, matches, len(matches), C.byref(P), C.c_float(scale), C.byref(res), C.byref(nm), tr))
    return _take_result(res, tr, n)


def mosaic_sequence(ctx, images, descs, kps, max_once, param=None, scale=1.0):
    """The chunk loop of MosaicUavVideo (M/MosaicWithoutPos.cpp:10252-10300) over decoded frames (uavm_mosaic_sequence).
    Returns a list of (first_frame, mosaic or None)."""
    n = len(images)
    ims, arr = _image_array(images)
    ds = [np.ascontiguousarray(d, np.float32) for d in descs]
    ks = [np.ascontiguousarray(k, np.float32) for k in kps]
    dp = (f32p * n)(*[_ptr(d, f32p) for d in ds]); kp = (f32p * n)(*[_ptr(k, f32p) for k in ks])
    nk = np.array([len(d) for d in ds], np.int32)
    P = _param(param)
    res = C.POINTER(L.Image)(); first = i32p(); nres = C.c_int(0)
    ctx.check(L.lib().uavm_mosaic_sequence(ctx._h, arr, n, dp, kp, _ptr(nk, i32p), C.byref(P), C.c_float(scale), int(max_once),
                                           C.byref(res), C.byref(first), C.byref(nres)))
    out = []
    for k in range(nres.value):
        r = res[k]
        img = None
        if r.imageData:
            buf = (C.c_uint8 * (r.widthStep * r.height)).from_address(r.imageData)
            img = np.frombuffer(buf, np.uint8).reshape(r.height, r.widthStep)[:, :r.width * 3].reshape(r.height, r.width, 3).copy()
            L.lib().uavm_free(C.c_void_p(r.imageData))
        out.append((int(first[k]), img))
    L.lib().uavm_free(C.cast(res, C.c_void_p)); L.lib().uavm_free(C.cast(first, C.c_void_p))
    return out


# ---- the reference's on-disk artefacts (host only; csrc/formats_host.cpp) --------------------------------------------------
def _io_check(rc, what, path):
    if rc != 0:
        raise UavmError(f"{what}({path!r}) failed with {rc}")


def read_match_file(path):
    """matchPairs.match (LoadMatchPairs, M/MosaicWithoutPos.cpp:4774-4797) -> ctypes array of MatchPointPairs."""
    n = C.c_int(0)
    _io_check(L.lib().uavm_match_file_count(path.encode(), C.byref(n)), "uavm_match_file_count", path)
    arr = (MatchPointPairs * max(n.value, 1))()
    _io_check(L.lib().uavm_match_file_read(path.encode(), arr, n.value, C.byref(n)), "uavm_match_file_read", path)
    return (MatchPointPairs * n.value).from_buffer(arr) if n.value else (MatchPointPairs * 0)()


def write_match_file(path, matches):
    _io_check(L.lib().uavm_match_file_write(path.encode(), matches, len(matches)), "uavm_match_file_write", path)


def write_match_text(path, matches):
    """matchPairs.txt (WriteMatchPairs_ASC2, :4751-4772)."""
    _io_check(L.lib().uavm_match_text_write(path.encode(), matches, len(matches)), "uavm_match_text_write", path)


def read_match_text(path, cap=1 << 22):
    arr = (MatchPointPairs * cap)(); n = C.c_int(0)
    _io_check(L.lib().uavm_match_text_read(path.encode(), arr, cap, C.byref(n)), "uavm_match_text_read", path)
    out = (MatchPointPairs * n.value)()
    C.memmove(out, arr, C.sizeof(MatchPointPairs) * n.value)
    return out


def _transforms_to_ctypes(T, fixed):
    T = np.ascontiguousarray(T, np.float32).reshape(-1, 9); n = len(T)
    arr = (ImageTransform * n)()
    for i in range(n):
        for j in range(9):
            arr[i].h.m[j] = float(T[i, j])
        arr[i].fixed = int(fixed[i])
    return arr


def write_transform_file(path, T, fixed):
    """tran0.txt (OutTransform, :2798-2818): rows for images 1..N-1."""
    arr = _transforms_to_ctypes(T, fixed)
    _io_check(L.lib().uavm_transform_file_write(path.encode(), arr, len(arr)), "uavm_transform_file_write", path)


def read_transform_file(path, cap=10000, imported=False):
    """tran0.txt -> (T (n,9) f32, fixed (n,)); imported=True reads ImportTransform's count-prefixed format (:2820-2843)."""
    arr = (ImageTransform * cap)(); n = C.c_int(0)
    fn = L.lib().uavm_transform_import if imported else L.lib().uavm_transform_file_read
    _io_check(fn(path.encode(), arr, cap, C.byref(n)), "uavm_transform_file_read", path)
    T = np.array([[arr[i].h.m[t] for t in range(9)] for i in range(n.value)], np.float32).reshape(-1, 9)
    return T, np.array([arr[i].fixed for i in range(n.value)], np.int32)


def write_feature_files(key_path, xml_path, keypoints, desc):
    """keypoint_%d.key + discriptor_%d.xml (WriteSurfKeyPoints, :4682-4704).  keypoints: ctypes array of _lib.KeyPoint."""
    desc = np.ascontiguousarray(desc, np.float32)
    _io_check(L.lib().uavm_key_file_write(key_path.encode(), keypoints, len(keypoints)), "uavm_key_file_write", key_path)
    _io_check(L.lib().uavm_descriptor_xml_write(xml_path.encode(), _ptr(desc, f32p), desc.shape[0], desc.shape[1]), "uavm_descriptor_xml_write", xml_path)


def read_feature_files(key_path, xml_path, cap=1 << 20):
    """LoadSurfKeyPoints (:4706-4734) -> (ctypes array of KeyPoint, (rows, cols) f32 descriptors)."""
    kp = (L.KeyPoint * cap)(); n = C.c_int(0)
    _io_check(L.lib().uavm_key_file_read(key_path.encode(), kp, cap, C.byref(n)), "uavm_key_file_read", key_path)
    out = (L.KeyPoint * n.value)()
    C.memmove(out, kp, C.sizeof(L.KeyPoint) * n.value)
    r = C.c_int(0); c = C.c_int(0)
    _io_check(L.lib().uavm_descriptor_xml_size(xml_path.encode(), C.byref(r), C.byref(c)), "uavm_descriptor_xml_size", xml_path)
    d = np.zeros((r.value, c.value), np.float32)
    _io_check(L.lib().uavm_descriptor_xml_read(xml_path.encode(), _ptr(d, f32p), d.size, C.byref(r), C.byref(c)), "uavm_descriptor_xml_read", xml_path)
    return out, d
