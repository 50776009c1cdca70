#!/usr/bin/env python
"""CUPTI timeline (torch.profiler) of two e2e steps of bench.py: prints the busy/idle structure of the H2D copies."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from torch.profiler import profile, ProfilerActivity

# reuse bench.run_ours' e2e_step by monkeypatching: simplest is to re-create the objects here
from imagemosaicing_b200 import api
W, H, NIMG, NKP = bench.W, bench.H, bench.NIMG, bench.NKP
descs, kps, Hs, T, base = bench.make_workload(0)
ctx = api.Context(0, torch.cuda.current_stream())
keep = np.ones(NIMG, np.int32); keep[0] = 0
cv = api.Canvas(ctx, T, W, H, keep)
h_frames = [torch.from_numpy(np.roll(base, (37 * k) % H, axis=0)).pin_memory() for k in range(NIMG)]
h_desc = [torch.from_numpy(d).pin_memory() for d in descs]; h_kp = [torch.from_numpy(k).pin_memory() for k in kps]
fs = api.FeatureSet(ctx, [NKP] * NIMG)
pairs = np.array([[i, i + 1] for i in range(NIMG - 1)], np.int32)
pb = api.PairBatch(ctx, fs, pairs)
GROUP = 7
def e2e_step():
    for k in range(NIMG): fs.upload(k, h_desc[k], h_kp[k])
    ctx.fork()
    pb.match(); pb.select(W, H); pb.ransac(2.5, 1000, base_seed=1000)
    ctx.unfork()
    for g0 in range(1, NIMG, GROUP):
        g1 = min(g0 + GROUP, NIMG)
        for k in range(g0, g1): cv.set_image(k, h_frames[k])
        cv.warp(g0, g1 - g0)
    ctx.join()
    return pb.collect(30)
for _ in range(2): e2e_step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(2): e2e_step()
    torch.cuda.synchronize()
os.makedirs("gpurun_out", exist_ok=True)
prof.export_chrome_trace("gpurun_out/e2e_trace.json")
ev = json.load(open("gpurun_out/e2e_trace.json"))["traceEvents"]
gpu = [e for e in ev if e.get("ph") == "X" and e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
gpu.sort(key=lambda e: e["ts"])
t0 = gpu[0]["ts"]
big = [e for e in gpu if e["cat"] == "gpu_memcpy" and e["dur"] > 300]
print("events", len(gpu), "big H2D copies", len(big), "span ms", (gpu[-1]["ts"] + gpu[-1]["dur"] - t0) / 1000)
prev_end = None
gaps = []
for e in big:
    if prev_end is not None and e["ts"] - prev_end > 20: gaps.append((round((e["ts"] - t0) / 1000, 2), round((e["ts"] - prev_end) / 1000, 3)))
    prev_end = e["ts"] + e["dur"]
print("gaps between consecutive frame copies (at ms, gap ms):", gaps[:40])
print("frame copy durations ms: min %.3f max %.3f" % (min(e["dur"] for e in big) / 1000, max(e["dur"] for e in big) / 1000))
# first/last events of each step
names = {}
for e in gpu:
    names.setdefault(e["name"][:40], []).append(((e["ts"] - t0) / 1000, e["dur"] / 1000))
for k, v in sorted(names.items(), key=lambda kv: kv[1][0][0]):
    print(f"{k:42s} n={len(v):4d} first@{v[0][0]:8.2f} last@{v[-1][0]:8.2f} total {sum(d for _, d in v):8.2f} ms")
