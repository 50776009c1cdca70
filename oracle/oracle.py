"""ctypes bindings for the CPU oracle (oracle/liboracle.so) and, when built, the compiled
reference RANSAC core (oracle/_ref/libref_ransac.so).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None

f32p = C.POINTER(C.c_float)
i32p = C.POINTER(C.c_int32)
u8p = C.POINTER(C.c_uint8)
f64p = C.POINTER(C.c_double)
i16p = C.POINTER(C.c_int16)


class RansacStats(C.Structure):
    _fields_ = [("best_tuple", C.c_int32), ("best_t", C.c_int32), ("max_support", C.c_int32),
                ("n_tuples", C.c_int32), ("n_counted", C.c_int32), ("n_refined", C.c_int32),
                ("n_inv_fail", C.c_int32), ("early_exit", C.c_int32), ("rand_calls", C.c_uint64)]


class ChipLayout(C.Structure):
    _fields_ = [("keep", C.c_int32), ("beg_x", C.c_int32), ("beg_y", C.c_int32),
                ("chip_w", C.c_int32), ("chip_h", C.c_int32), ("sx", C.c_float), ("sy", C.c_float),
                ("quad", C.c_float * 8), ("inv", C.c_float * 9)]


class CanvasLayout(C.Structure):
    _fields_ = [("canvas_w", C.c_int32), ("canvas_h", C.c_int32), ("dgx", C.c_float), ("dgy", C.c_float)]


def build(force=False):
    """Compile liboracle.so (and _ref/libref_ransac.so when /root/reference is present)."""
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("oracle.c", "oracle_blend.c", "oracle_fast.c", "oracle.h") if os.path.exists(os.path.join(_HERE, f))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"], stdout=subprocess.DEVNULL)
    ref_so = os.path.join(_HERE, "_ref", "libref_ransac.so")
    if os.path.isdir("/root/reference/code/MosaicingCode/mosaicing") and (force or not os.path.exists(ref_so)):
        subprocess.check_call(["make", "-C", _HERE, "-s", "ref"], stdout=subprocess.DEVNULL)


def lib():
    global _LIB
    if _LIB is None:
        build()
        _LIB = C.CDLL(os.environ.get("ORACLE_LIB_PATH") or os.path.join(_HERE, "liboracle.so"))   # ORACLE_LIB_PATH: sanitizer build
        _LIB.orc_ransac2d.restype = C.c_int
        _LIB.orc_select.restype = C.c_int
    return _LIB


def ref():
    """The reference's own Ransac2D core compiled from /root/reference; None if not built."""
    global _REF
    if _REF is None:
        p = os.path.join(_HERE, "_ref", "libref_ransac.so")
        if not os.path.exists(p):
            try:
                build()
            except Exception:
                pass
        if not os.path.exists(p):
            return None
        _REF = C.CDLL(p)
    return _REF


def set_threads(n):
    return lib().orc_set_threads(int(n))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, t):
    return a.ctypes.data_as(t)


# ----------------------------------------------------------------------------------------------
def match_l2(A, B):
    A = np.ascontiguousarray(A, dtype=np.uint8); B = np.ascontiguousarray(B, dtype=np.uint8)
    na, dim = A.shape; nb = B.shape[0]
    idx = np.empty(na, np.int32); d2 = np.empty(na, np.int32)
    lib().orc_match_l2(_p(A, u8p), na, _p(B, u8p), nb, dim, _p(idx, i32p), _p(d2, i32p))
    return idx, d2


def match_l2_fast(A, B):
    """Same result as match_l2 (exact, lowest index on ties), tuned (AVX2 integer dot products): bench.py's CPU arm."""
    A = np.ascontiguousarray(A, dtype=np.uint8); B = np.ascontiguousarray(B, dtype=np.uint8)
    na, dim = A.shape; nb = B.shape[0]
    idx = np.empty(na, np.int32); d2 = np.empty(na, np.int32)
    lib().orc_match_l2_fast(_p(A, u8p), na, _p(B, u8p), nb, dim, _p(idx, i32p), _p(d2, i32p))
    return idx, d2


def select(train_idx, d2, kp1, kp2, width, height, grid_x=3, grid_y=3, max_num=400, frac=0.3):
    train_idx = np.ascontiguousarray(train_idx, np.int32); d2 = np.ascontiguousarray(d2, np.int32)
    kp1 = _f32(kp1); kp2 = _f32(kp2)
    n = len(train_idx)
    cap = 2048
    xy1 = np.zeros((cap, 2), np.float32); xy2 = np.zeros((cap, 2), np.float32)
    id1 = np.zeros(cap, np.int32); id2 = np.zeros(cap, np.int32)
    cnt = lib().orc_select(_p(train_idx, i32p), _p(d2, i32p), n, _p(kp1, f32p), _p(kp2, f32p),
                           width, height, grid_x, grid_y, max_num, C.c_double(frac),
                           _p(xy1, f32p), _p(id1, i32p), _p(xy2, f32p), _p(id2, i32p))
    return xy1[:cnt].copy(), id1[:cnt].copy(), xy2[:cnt].copy(), id2[:cnt].copy()


def ransac2d(xy1, xy2, dist=2.5, sample_times=1000, seed=1):
    xy1 = _f32(xy1); xy2 = _f32(xy2); n = len(xy1)
    mask = np.zeros(max(n, 1), np.uint8); H = np.zeros(9, np.float32); ni = C.c_int(0); st = RansacStats()
    ok = lib().orc_ransac2d(_p(xy1, f32p), _p(xy2, f32p), n, C.c_float(dist), sample_times,
                            C.c_uint32(seed), _p(mask, u8p), _p(H, f32p), C.byref(ni), C.byref(st))
    return ok, mask[:n].copy(), H, ni.value, st


def ransac_eval_tuple(xy1, xy2, idx, thr2=6.25):
    xy1 = _f32(xy1); xy2 = _f32(xy2); idx = np.ascontiguousarray(idx, np.int32)
    h = np.zeros(9, np.float32); sup = C.c_int(0)
    kind = lib().orc_ransac_eval_tuple(_p(xy1, f32p), _p(xy2, f32p), len(xy1), _p(idx, i32p),
                                       C.c_float(thr2), _p(h, f32p), C.byref(sup))
    return kind, h, sup.value


def solve_homography(xy1, xy2, which="oracle"):
    xy1 = _f32(xy1); xy2 = _f32(xy2); h = np.zeros(9, np.float32)
    fn = lib().orc_solve_homography if which == "oracle" else ref().ref_solve_homography
    fn(_p(xy1, f32p), _p(xy2, f32p), len(xy1), _p(h, f32p))
    return h


def nls_projection2(xy1, xy2, init, stop=1e-10, which="oracle"):
    xy1 = _f32(xy1); xy2 = _f32(xy2); init = _f32(init); out = np.zeros(9, np.float32)
    if which == "oracle":
        lib().orc_nls_projection2(_p(xy1, f32p), _p(xy2, f32p), len(xy1), _p(out, f32p), _p(init, f32p),
                                  C.c_float(stop), None)
    else:
        ref().ref_nls_projection2(_p(xy1, f32p), _p(xy2, f32p), len(xy1), _p(out, f32p), _p(init, f32p),
                                  C.c_float(stop))
    return out


def inverse_matrix(src, eps, which="oracle"):
    src = _f32(src); n = src.shape[0]; dst = np.zeros((n, n), np.float32)
    fn = lib().orc_inverse_matrix if which == "oracle" else ref().ref_inverse_matrix
    rc = fn(_p(src, f32p), n, _p(dst, f32p), C.c_float(eps))
    return rc, dst


def ref_ransac2d(xy1, xy2, dist=2.5, sample_times=1000, seed=1):
    xy1 = _f32(xy1); xy2 = _f32(xy2); n = len(xy1)
    mask = np.zeros(max(n, 1), np.uint8); H = np.zeros(9, np.float32); ni = C.c_int(0); rc = C.c_uint64(0)
    ok = ref().ref_ransac2d(_p(xy1, f32p), _p(xy2, f32p), n, C.c_float(dist), sample_times,
                            C.c_uint32(seed), _p(mask, u8p), _p(H, f32p), C.byref(ni), C.byref(rc))
    return ok, mask[:n].copy(), H, ni.value, rc.value


def align_affine(pairs, fixed_img, T0=None):
    """pairs: (P, 8) [imgA xA yA fixedA imgB xB yB fixedB]; fixed_img: (N,) 0/1."""
    pairs = np.ascontiguousarray(pairs, np.float64); fixed_img = np.ascontiguousarray(fixed_img, np.int32)
    n = len(fixed_img)
    if T0 is None:
        T0 = np.tile(np.eye(3, dtype=np.float32).reshape(1, 9), (n, 1))
    T0 = _f32(T0); out = np.zeros((n, 9), np.float32)
    rc = lib().orc_align_affine(_p(pairs, f64p), len(pairs), _p(fixed_img, i32p), _p(T0, f32p), n, _p(out, f32p))
    return rc, out


def canvas_layout(H, keep, img_w, img_h):
    H = _f32(H).reshape(-1, 9); n = len(H)
    keep = np.ascontiguousarray(keep if keep is not None else np.ones(n), np.int32)
    canvas = CanvasLayout(); chips = (ChipLayout * n)()
    lib().orc_canvas_layout_compute(_p(H, f32p), _p(keep, i32p), n, img_w, img_h, C.byref(canvas), chips)
    return canvas, chips


def warp_chip(img, canvas, chip):
    img = np.ascontiguousarray(img, np.uint8); h, w = img.shape[:2]
    out = np.zeros((chip.chip_h, chip.chip_w, 3), np.uint8); mask = np.zeros((chip.chip_h, chip.chip_w), np.uint8)
    lib().orc_warp_chip(_p(img, u8p), w, h, img.strides[0], C.byref(canvas), C.byref(chip),
                        _p(out, u8p), out.strides[0], _p(mask, u8p), mask.strides[0])
    return out, mask


def seam_masks(masks, chips_valid, canvas_w, canvas_h):
    """masks: list of (h, w) u8 arrays (modified copies returned); chips_valid: list of ChipLayout."""
    n = len(masks)
    ms = [np.ascontiguousarray(m, np.uint8).copy() for m in masks]
    ptrs = (u8p * n)(*[_p(m, u8p) for m in ms])
    steps = np.array([m.strides[0] for m in ms], np.int32)
    arr = (ChipLayout * n)(*chips_valid)
    lib().orc_seam_masks(ptrs, _p(steps, i32p), arr, n, canvas_w, canvas_h)
    return ms


# ---- the reference's own warp loop / FindMasksByDistMap compiled from /root/reference (oracle/_ref) ----
def ref_warp_chip(img, canvas, chip):
    """M/MosaicImage.cpp:2350-2448 compiled in place; chip bytes start at zero (the reference leaves
    invalid pixels uninitialised)."""
    img = np.ascontiguousarray(img, np.uint8); h, w = img.shape[:2]
    out = np.zeros((chip.chip_h, chip.chip_w, 3), np.uint8); mask = np.zeros((chip.chip_h, chip.chip_w), np.uint8)
    inv = np.array(list(chip.inv), np.float32)
    ref().ref_warp_chip(_p(img, u8p), w, h, img.strides[0], C.c_float(canvas.dgx), C.c_float(canvas.dgy),
                        C.c_float(chip.sx), C.c_float(chip.sy), chip.beg_x, chip.beg_y, _p(inv, f32p),
                        chip.chip_w, chip.chip_h, _p(out, u8p), out.strides[0], _p(mask, u8p), mask.strides[0])
    return out, mask


def ref_seam_masks(masks, chips_valid, canvas_w, canvas_h):
    n = len(masks)
    ms = [np.ascontiguousarray(m, np.uint8).copy() for m in masks]
    ptrs = (u8p * n)(*[_p(m, u8p) for m in ms])
    steps = np.array([m.strides[0] for m in ms], np.int32)
    cw = np.array([c.chip_w for c in chips_valid], np.int32); chh = np.array([c.chip_h for c in chips_valid], np.int32)
    quads = np.array([list(c.quad) for c in chips_valid], np.float32)
    tl = np.array([[c.beg_x, c.beg_y] for c in chips_valid], np.int32)
    ref().ref_find_masks(ptrs, _p(steps, i32p), _p(cw, i32p), _p(chh, i32p), _p(quads, f32p), _p(tl, i32p), n, canvas_w, canvas_h)
    return ms


# ---- K7: multi-band blend ---------------------------------------------------------------------------
def pyr_down_s16(a):
    a = np.ascontiguousarray(a, np.int16); h, w = a.shape[:2]; ch = 1 if a.ndim == 2 else a.shape[2]
    out = np.zeros(((h + 1) // 2, (w + 1) // 2) + (() if a.ndim == 2 else (ch,)), np.int16)
    lib().orc_pyr_down_s16(_p(a, i16p), w, h, ch, _p(out, i16p))
    return out


def pyr_up_s16(a):
    a = np.ascontiguousarray(a, np.int16); h, w = a.shape[:2]; ch = 1 if a.ndim == 2 else a.shape[2]
    out = np.zeros((2 * h, 2 * w) + (() if a.ndim == 2 else (ch,)), np.int16)
    lib().orc_pyr_up_s16(_p(a, i16p), w, h, ch, _p(out, i16p), 2 * w, 2 * h)
    return out


def pyr_down_f32(a):
    a = np.ascontiguousarray(a, np.float32); h, w = a.shape
    out = np.zeros(((h + 1) // 2, (w + 1) // 2), np.float32)
    lib().orc_pyr_down_f32(_p(a, f32p), w, h, _p(out, f32p))
    return out


def multiband_blend(chips_u8, masks, tls, canvas_w, canvas_h, num_bands=5):
    """chips_u8: list of (h,w,3) u8 (converted to int16 like Mat::convertTo(CV_16S)); masks: list of (h,w) u8;
    tls: list of (x, y).  Returns (canvas u8 (H,W,3), mask u8 (H,W))."""
    n = len(chips_u8)
    c16 = [np.ascontiguousarray(c, np.uint8).astype(np.int16) for c in chips_u8]
    ms = [np.ascontiguousarray(m, np.uint8) for m in masks]
    cp = (i16p * n)(*[_p(c, i16p) for c in c16]); mp = (u8p * n)(*[_p(m, u8p) for m in ms])
    tx = np.array([t[0] for t in tls], np.int32); ty = np.array([t[1] for t in tls], np.int32)
    cw = np.array([c.shape[1] for c in c16], np.int32); chh = np.array([c.shape[0] for c in c16], np.int32)
    out = np.zeros((canvas_h, canvas_w, 3), np.uint8); om = np.zeros((canvas_h, canvas_w), np.uint8)
    lib().orc_multiband_blend(cp, mp, _p(tx, i32p), _p(ty, i32p), _p(cw, i32p), _p(chh, i32p), n, canvas_w, canvas_h,
                              num_bands, _p(out, u8p), _p(om, u8p))
    return out, om


def paste(H, imgs):
    """MosaicImagesRefined (blending != 2): returns the canvas (h, w, 3) u8."""
    H = _f32(H).reshape(-1, 9); n = len(H)
    ims = [np.ascontiguousarray(i, np.uint8) for i in imgs]
    h, w = ims[0].shape[:2]
    ptrs = (u8p * n)(*[_p(i, u8p) for i in ims])
    ow = C.c_int(0); oh = C.c_int(0)
    lib().orc_paste(_p(H, f32p), n, w, h, ptrs, ims[0].strides[0], C.byref(ow), C.byref(oh), None, 0)
    out = np.zeros((oh.value, ow.value, 3), np.uint8)
    lib().orc_paste(_p(H, f32p), n, w, h, ptrs, ims[0].strides[0], C.byref(ow), C.byref(oh), _p(out, u8p), out.strides[0])
    return out
