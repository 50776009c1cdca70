#!/usr/bin/env python
"""Times K6 (seam masks) and K7 (multi-band blend) on the bench strip and prints a checksum of the mosaic."""
import hashlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from imagemosaicing_b200 import api, synth
n_img = int(os.environ.get("AB_NIMG", "50"))
rng = np.random.default_rng(5)
_, _, Hs = synth.make_strip(n_img, bench.W, bench.H, 64, seed=synth.SEED_BASE)
T = bench.strip_transforms(Hs)
base = synth.texture_image(rng, bench.W, bench.H, 6)
ctx = api.Context(0, torch.cuda.current_stream())
keep = np.ones(n_img, np.int32); keep[0] = 0
cv = api.Canvas(ctx, T, bench.W, bench.H, keep)
for k in range(1, n_img):
    cv.set_image(k, np.roll(base, (37 * k) % bench.H, axis=0))
cv.warp(); cv.seam_masks(); cv.blend(5); torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
cv.warp()
e[0].record(); cv.seam_masks(); e[1].record(); cv.blend(5); e[2].record(); torch.cuda.synchronize()
out, m = cv.result()
h = hashlib.sha256(); h.update(out[::7].tobytes()); h.update(m[::7].tobytes())
print(f"k6_ms={e[0].elapsed_time(e[1]):.3f} k7_ms={e[1].elapsed_time(e[2]):.3f} canvas={out.shape} sha={h.hexdigest()[:16]}")
if os.environ.get("AB_TRACE"):
    from torch.profiler import profile, ProfilerActivity
    cv.warp(); torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        cv.seam_masks(); cv.blend(5); torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=12, max_name_column_width=60))
