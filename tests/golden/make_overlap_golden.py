"""Generates tests/golden/overlap_golden.npz: transforms + keep flags from the REFERENCE's own
ResampleByOverlap (M/MosaicImage.cpp:2070-2201), compiled in place from /root/reference
(`make -C oracle ref`).  Run in the build container; the .npz travels to the GPU box.

    python tests/golden/make_overlap_golden.py
"""
import ctypes as C
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O            # noqa: E402
from imagemosaicing_b200 import synth    # noqa: E402

R = O.ref()
assert R is not None
f32p = C.POINTER(C.c_float); i32p = C.POINTER(C.c_int32)
rng = np.random.default_rng(20160308)
out = {}
k = 0
for trial in range(60):
    n = int(rng.integers(2, 14)); w, h = [(1000, 750), (4000, 3000)][trial % 2]
    T = [np.eye(3)]
    for i in range(1, n):
        Hk = synth.pair_homography(rng, w, h, overlap=(0.7, 0.98))
        if trial % 3:
            Hk[2, :2] = 0
        T.append(T[-1] @ Hk)
    Hs = np.stack([t / t[2, 2] for t in T]).astype(np.float32).reshape(n, 9)
    if trial % 7 == 0:
        Hs[rng.integers(0, n), 8] = 0
    keep = np.zeros(n, np.int32)
    R.ref_resample_by_overlap(Hs.ctypes.data_as(f32p), n, w, h, C.c_float(0.7), keep.ctypes.data_as(i32p))
    out[f"H_{k}"] = Hs; out[f"keep_{k}"] = keep; out[f"wh_{k}"] = np.array([w, h]); k += 1
out["n_cases"] = np.array([k])
p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "overlap_golden.npz")
np.savez_compressed(p, **out)
print("wrote", p, k, "cases; dropped", sum(int((out[f'keep_{i}'] == 0).sum()) for i in range(k)))
