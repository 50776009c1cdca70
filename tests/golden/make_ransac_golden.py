"""Generates tests/golden/ransac_golden.npz: inputs + outputs of the REFERENCE's own Ransac2D
(compiled in place from /root/reference by `make -C oracle ref`, sample stream = MSVC LCG with the
given seed).  Run in the build container (where /root/reference exists); the .npz travels to the GPU box.

    python tests/golden/make_ransac_golden.py
"""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O            # noqa: E402
from imagemosaicing_b200 import synth    # noqa: E402

assert O.ref() is not None, "oracle/_ref/libref_ransac.so missing (needs /root/reference)"
rng = np.random.default_rng(20160308)
cases = {}
k = 0
for w, h in [(4000, 3000), (1000, 750), (512, 512)]:
    for n in (4, 9, 57, 200, 396):
        for inl in (0.3, 0.6, 1.0):
            xy1, xy2, _ = synth.make_candidates(rng, n, w, h, inl, 0.5 if inl < 1.0 else 0.0)
            seed = int(rng.integers(0, 2 ** 32))
            ok, mask, H, ninl, rc = O.ref_ransac2d(xy1, xy2, 2.5, 1000, seed)
            cases[f"xy1_{k}"] = xy1; cases[f"xy2_{k}"] = xy2
            cases[f"meta_{k}"] = np.array([seed, ok, ninl, rc], np.int64)
            cases[f"mask_{k}"] = mask; cases[f"H_{k}"] = H
            k += 1
cases["n_cases"] = np.array([k])
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ransac_golden.npz")
np.savez_compressed(out, **cases)
print("wrote", out, k, "cases", os.path.getsize(out), "bytes")
