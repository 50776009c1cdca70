"""Generates tests/golden/warp_golden.npz: small frames, transforms, and the chips / validity masks / seam masks
produced by the REFERENCE's own warp loop (M/MosaicImage.cpp:2350-2448) and FindMasksByDistMap (:1761-1881),
compiled in place from /root/reference (`make -C oracle ref`).  Run in the build container.

    python tests/golden/make_warp_golden.py
"""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O            # noqa: E402
from imagemosaicing_b200 import synth    # noqa: E402

assert O.ref() is not None
rng = np.random.default_rng(20160308)
out = {}
cases = [(160, 120, 3, False), (200, 150, 4, True), (97, 131, 2, False)]
for ci, (w, h, n, proj) in enumerate(cases):
    T = [np.eye(3)]
    for k in range(1, n):
        Hk = synth.pair_homography(rng, w, h, overlap=(0.5, 0.8))
        if not proj:
            Hk[2, :2] = 0
        T.append(T[-1] @ Hk)
    H = np.stack([t / t[2, 2] for t in T]).astype(np.float32).reshape(n, 9)
    imgs = [synth.texture_image(rng, w, h, 5) for _ in range(n)]
    canvas, chips = O.canvas_layout(H, None, w, h)
    masks = []
    out[f"H_{ci}"] = H; out[f"whn_{ci}"] = np.array([w, h, n])
    for k in range(n):
        px, m = O.ref_warp_chip(imgs[k], canvas, chips[k])
        out[f"img_{ci}_{k}"] = imgs[k]; out[f"chip_{ci}_{k}"] = px; out[f"mask_{ci}_{k}"] = m
        masks.append(m)
    seam = O.ref_seam_masks(masks, [chips[k] for k in range(n)], canvas.canvas_w, canvas.canvas_h)
    for k in range(n):
        out[f"seam_{ci}_{k}"] = seam[k]
out["n_cases"] = np.array([len(cases)])
p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "warp_golden.npz")
np.savez_compressed(p, **out)
print("wrote", p, os.path.getsize(p), "bytes")
