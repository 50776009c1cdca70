"""imagemosaicing_b200 — B200-native (sm_100a) hot path of UAV image mosaicking:
match -> select -> RANSAC homography -> global affine alignment -> warp / seam masks / multi-band blend.
The compute lives in libuavmosaic.so (hand-written CUDA behind the C ABI of include/uavm.h);
this package is the Python host mirror of the reference's entry points for that path."""
from . import _lib  # noqa: F401

__all__ = ["_lib", "api", "synth"]
