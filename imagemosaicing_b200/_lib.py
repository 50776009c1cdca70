"""Loader for the in-tree C-ABI library (imagemosaicing_b200/libuavmosaic.so).

There is no CPU or PyTorch fallback: if the library is missing this module raises, and every
compute entry point returns -2 when no sm_100 GPU is present.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("UAVM_LIB_PATH") or os.path.join(_HERE, "libuavmosaic.so")   # UAVM_LIB_PATH: sanitizer / debug builds
_lib = None

f32p = C.POINTER(C.c_float)
i32p = C.POINTER(C.c_int32)
u32p = C.POINTER(C.c_uint32)
u8p = C.POINTER(C.c_uint8)


class SfPoint(C.Structure):          # M/Point.h:27-47
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("id", C.c_int32)]


class DMatch(C.Structure):           # cv::DMatch
    _fields_ = [("queryIdx", C.c_int32), ("trainIdx", C.c_int32), ("imgIdx", C.c_int32), ("distance", C.c_float)]


class ProjectMat(C.Structure):       # M/Bitmap.h:42-45
    _fields_ = [("m", C.c_float * 9)]


class ImageTransform(C.Structure):   # M/MosaicWithoutPos.h:224-228
    _fields_ = [("h", ProjectMat), ("fixed", C.c_int32)]


class MatchPointPairs(C.Structure):  # M/MosaicWithoutPos.h:135-153 (40 bytes)
    _fields_ = [("ptA", SfPoint), ("ptA_i", C.c_int32), ("ptA_Fixed", C.c_int32),
                ("ptB", SfPoint), ("ptB_i", C.c_int32), ("ptB_Fixed", C.c_int32)]


class KeyPoint(C.Structure):          # cv::KeyPoint of OpenCV 2.4 (28 bytes), record of keypoint_%d.key
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("size", C.c_float), ("angle", C.c_float), ("response", C.c_float),
                ("octave", C.c_int32), ("class_id", C.c_int32)]


class Image(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("nChannels", C.c_int32), ("widthStep", C.c_int32),
                ("imageData", C.c_void_p)]


class Param(C.Structure):
    _fields_ = [("ransacDist", C.c_float), ("blending", C.c_int32), ("loadMatchPairs", C.c_int32),
                ("sampleTimes", C.c_int32), ("pairWindow", C.c_int32), ("minInnerPoints", C.c_int32),
                ("gridX", C.c_int32), ("gridY", C.c_int32), ("maxNum", C.c_int32), ("matchFrac", C.c_float),
                ("numBands", C.c_int32), ("overlapT", C.c_float), ("seed", C.c_uint32)]


class RansacResult(C.Structure):
    _fields_ = [("ok", C.c_int32), ("n_inliers", C.c_int32), ("max_support", C.c_int32), ("best_tuple", C.c_int32),
                ("n_tuples", C.c_int32), ("n_counted", C.c_int32), ("H", C.c_float * 9)]


class ChipLayout(C.Structure):
    _fields_ = [("keep", C.c_int32), ("beg_x", C.c_int32), ("beg_y", C.c_int32), ("chip_w", C.c_int32),
                ("chip_h", C.c_int32), ("sx", C.c_float), ("sy", C.c_float), ("quad", C.c_float * 8),
                ("inv", C.c_float * 9)]


class CanvasLayout(C.Structure):
    _fields_ = [("canvas_w", C.c_int32), ("canvas_h", C.c_int32), ("dgx", C.c_float), ("dgy", C.c_float)]


def build(verbose=False):
    """Compile the library in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
    cmd = ["make", "-C", os.path.join(_HERE, "csrc"), "-j8"]
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout)
    if out.returncode != 0:
        raise RuntimeError("building libuavmosaic.so failed")
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        try:                                   # absent only in a host-only sanitizer build (UAVM_LIB_PATH)
            L.uavm_last_error.restype = C.c_char_p
            L.uavm_ctx_launch_count.restype = C.c_int64
            L.uavm_jpeg_decode_bgr.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_int]
            L.uavm_canvas_set_image_jpeg.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int64]
            L.uavm_canvas_set_images_jpeg.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
            L.uavm_jpeg_hw_engines.argtypes = [C.c_void_p]
            L.uavm_jpeg_set_threads.argtypes = [C.c_void_p, C.c_int]
            L.uavm_canvas_bind_root.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
            L.uavm_canvas_bound_root.argtypes = [C.c_void_p]
            L.uavm_jpeg_destroy.restype = None
            L.uavm_jpeg_destroy.argtypes = [C.c_void_p, C.c_void_p]
            L.uavm_dist_destroy.restype = None
            L.uavm_dist_destroy.argtypes = [C.c_void_p, C.c_void_p]
            L.uavm_pairbatch_allgather.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
            L.uavm_canvas_gather.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
            L.uavm_dist_broadcast.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int]
        except AttributeError:
            if not os.environ.get("UAVM_LIB_PATH"):
                raise
        _lib = L
    return _lib
