#!/usr/bin/env python
"""BASELINE configs[3] microbench: 32768 x 32768 SIFT-128 L2 1-NN on the tcgen05 path (k2_match_tcgen05).
Prints ms per launch and the algorithmic TFLOP/s (2 * 32768^2 * 128 per launch) against the measured bf16 peak."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from imagemosaicing_b200 import api, synth
import bench
n = 32768
rng = np.random.default_rng(4)
A = synth.sift_like_descriptors(rng, n); B = synth.sift_like_descriptors(rng, n)
ctx = api.Context(0, torch.cuda.current_stream())
fs = api.FeatureSet(ctx, [n, n]); fs.upload(0, A, None); fs.upload(1, B, None)
pb = api.PairBatch(ctx, fs, [[0, 1]])
for _ in range(3): pb.match()
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
k = 20
e0.record()
for _ in range(k): pb.match()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / k
tf = 2.0 * n * n * 128 / (ms * 1e-3) / 1e12
pk = bench.peaks()
print(json.dumps({"workload": "configs[3]: 32768 x 32768 x 128 u8 distance matrix + fused arg-min", "ms_per_launch": ms, "achieved_tflops": tf,
                  "peak_bf16_tflops": pk["bf16_tflops"], "frac_of_bf16_peak": tf / pk["bf16_tflops"], "peak_src": pk["src"],
                  "compulsory_bytes": 2 * n * 128 + n * 8, "distance_matrix_bytes_never_written": 4 * n * n}))
