#!/usr/bin/env python
"""Where does the end-to-end step go?  Times (a) the 49 frame uploads alone through Canvas.set_image, (b) + warp groups,
(c) the descriptor uploads alone, on the bench workload."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from imagemosaicing_b200 import api, synth
W, H, NIMG, NKP = bench.W, bench.H, bench.NIMG, bench.NKP
descs, kps, Hs, T, base = bench.make_workload(0)
ctx = api.Context(0, torch.cuda.current_stream())
keep = np.ones(NIMG, np.int32); keep[0] = 0
cv = api.Canvas(ctx, T, W, H, keep)
h_frames = [torch.from_numpy(np.roll(base, (37 * k) % H, axis=0)).pin_memory() for k in range(NIMG)]
h_desc = [torch.from_numpy(d).pin_memory() for d in descs]; h_kp = [torch.from_numpy(k).pin_memory() for k in kps]
fs = api.FeatureSet(ctx, [NKP] * NIMG)
def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (time.perf_counter() - t0) * 1000 / n
def frames():
    for k in range(1, NIMG): cv.set_image(k, h_frames[k])
def frames_warp():
    for g0 in range(1, NIMG, 7):
        for k in range(g0, min(g0 + 7, NIMG)): cv.set_image(k, h_frames[k])
        cv.warp(g0, min(7, NIMG - g0))
def feats():
    for k in range(NIMG): fs.upload(k, h_desc[k], h_kp[k])
for name, fn in [("frames", frames), ("frames+warp", frames_warp), ("features", feats)]:
    g, w = timeit(fn)
    print(f"{name:12s} gpu {g:7.2f} ms  wall {w:7.2f} ms   ({(NIMG-1)*W*H*3/g/1e6 if 'frames' in name else NIMG*NKP*136/g/1e6:.1f} GB/s)")
