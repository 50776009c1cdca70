#!/usr/bin/env python
"""A/B timing of the K5 warp kernels on the bench workload (49 chips of 4000x3000).  Run twice:
   python scripts/ab_warp.py            (packed f32x2 kernel)
   UAVM_K5_SCALAR=1 python scripts/ab_warp.py   (scalar kernel)
Prints ms per launch and a checksum of all chips + masks (both variants must print the same checksum)."""
import hashlib
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from imagemosaicing_b200 import api

n_img = int(os.environ.get("AB_NIMG", "50"))
from imagemosaicing_b200 import synth
rng = np.random.default_rng(5)
_, _, Hs = synth.make_strip(n_img, bench.W, bench.H, 64, seed=synth.SEED_BASE)
T = bench.strip_transforms(Hs)
base = synth.texture_image(rng, bench.W, bench.H, 6)
ctx = api.Context(0, torch.cuda.current_stream())
keep = np.ones(n_img, np.int32); keep[0] = 0
cv = api.Canvas(ctx, T, bench.W, bench.H, keep)
for k in range(1, n_img):
    cv.set_image(k, np.roll(base, (37 * k) % bench.H, axis=0))
for _ in range(3):
    cv.warp()
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
n = 10
e0.record()
for _ in range(n):
    cv.warp()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
h = hashlib.sha256()
for k in range(1, n_img, 7):
    px, m = cv.chip(k)
    h.update(px.tobytes()); h.update(m.tobytes())
print(f"variant={'scalar' if os.environ.get('UAVM_K5_SCALAR') else 'packed'} ms_per_launch={ms:.4f} sha={h.hexdigest()[:16]}")
