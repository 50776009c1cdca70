#!/usr/bin/env python
"""BASELINE configs[2]: 200-image UAV block (10 strips x 20 frames of 4000x3000, 8192 keypoints each, shared world-point
model), ALL pairs whose footprints overlap -> GPU match / select / RANSAC sharded over the ranks (no data-path collective),
one gather of the accepted inlier matches, connectivity + global affine alignment on rank 0, accuracy against the
ground-truth poses.

  python scripts/bench_block200.py [--rows 10 --cols 20]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_block200.py
"""
import argparse, ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.distributed as dist
import bench
from imagemosaicing_b200 import api, synth, dist as D, _lib as L

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=10); ap.add_argument("--cols", type=int, default=20)
ap.add_argument("--kp", type=int, default=8192); ap.add_argument("--steps", type=int, default=5)
args = ap.parse_args()
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
W, H = bench.W, bench.H
n = args.rows * args.cols
t0 = time.perf_counter()
descs, kps, poses, pairs = synth.make_block(args.rows, args.cols, W, H, args.kp, seed=synth.SEED_BASE + 3)
t_synth = time.perf_counter() - t0
ctx = api.Context(local, torch.cuda.current_stream())
fs = api.FeatureSet(ctx, [args.kp] * n)
for i in range(n):
    fs.upload(i, descs[i], kps[i])                      # descriptors replicated on every rank (210 MB)
mine = D.shard_pairs(len(pairs), rank, world)
pb = api.PairBatch(ctx, fs, pairs[mine])
seeds = (1000 + mine).astype(np.uint32)

def pair_path():
    pb.match(); pb.select(W, H); pb.ransac(2.5, 1000, seeds=seeds)
for _ in range(2): pair_path()
torch.cuda.synchronize()
if world > 1: dist.barrier()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps): pair_path()
e1.record(); torch.cuda.synchronize()
pair_ms = e0.elapsed_time(e1) / args.steps
# accepted matches of this rank, merged over the ranks in global pair order (two all-reduces of disjoint slots)
out, n_m, n_acc = pb.collect(30)
rec = np.frombuffer(bytes(out)[:n_m * 40], dtype=D.MPP_DTYPE).copy()
t1 = time.perf_counter()
rec_all = D.gather_match_pairs(D.split_collected(rec, pairs[mine]), mine, len(pairs), device=dev)     # global pair order, identical on every rank
gather_ms = (time.perf_counter() - t1) * 1e3
tt = torch.tensor([pair_ms], device=dev, dtype=torch.float64)
if world > 1: dist.all_reduce(tt, op=dist.ReduceOp.MAX)
if rank == 0:
    n_all = len(rec_all)
    mp = (L.MatchPointPairs * n_all).from_buffer_copy(rec_all.tobytes())
    ta = time.perf_counter()
    label = (C.c_int32 * n)()
    L.lib().uavm_connected_images(mp, n_all, n, label)
    for r in mp:
        if r.ptA_i == 0: r.ptA_Fixed = 1
        if r.ptB_i == 0: r.ptB_Fixed = 1
    init = (L.ImageTransform * n)(); res = (L.ImageTransform * n)()
    nfix = 0
    for i in range(n):
        for t in range(9): init[i].h.m[t] = 1.0 if t in (0, 4, 8) else 0.0
        init[i].fixed = 1 if (i == 0 or label[i] == 0) else 0; nfix += init[i].fixed
    keep = [r for r in mp if label[r.ptA_i] and label[r.ptB_i]]
    mp2 = (L.MatchPointPairs * len(keep))(*keep)
    tb = time.perf_counter()
    rc = L.lib().uavm_align_affine(mp2, len(keep), init, n, nfix, res)
    solve_ms = (time.perf_counter() - tb) * 1e3
    align_ms = (time.perf_counter() - ta) * 1e3
    corners = np.array([[0, 0], [W - 1, 0], [W - 1, H - 1], [0, H - 1]], np.float64)
    err = []
    for k in range(n):
        if not label[k]: continue
        G = np.linalg.inv(poses[0]) @ poses[k]
        T = np.array([res[k].h.m[t] for t in range(9)], np.float64).reshape(3, 3)
        err.append(float(np.abs(synth.apply_h(G, corners) - synth.apply_h(T, corners)).max()))
    acc_pairs = len(set((r.ptA_i, r.ptB_i) for r in mp))
    print(json.dumps({"workload": f"configs[2]: {n}-image block ({args.rows} strips x {args.cols}), {W}x{H}, {args.kp} kp/image, all {len(pairs)} pairs in overlap",
                      "n_gpus": world, "pairs": int(len(pairs)), "accepted_pairs": acc_pairs, "inlier_matches": n_all,
                      "pair_path_ms": float(tt[0]), "pairs_per_s_match_select_ransac": len(pairs) / (float(tt[0]) / 1e3),
                      "gather_ms_host_clock": gather_ms, "connectivity_plus_alignment_ms_host": align_ms, "uavm_align_affine_ms": solve_ms, "align_rc": rc,
                      "unknowns": 6 * (n - nfix), "connected_images": int(sum(label)),
                      "corner_error_px_max": max(err), "corner_error_px_median": float(np.median(err)), "synth_s": t_synth}))
if world > 1: dist.destroy_process_group()
