#!/usr/bin/env python
"""BASELINE configs[4] microbench: cols x rows warped tiles (default 25 x 20 = 500 frames of 4000x3000 on a (1500, 1947) px
pitch with +-2 deg / +-2 % jitter) -> one ~40000 x 40000 mosaic canvas: K5 warp + K6 seam masks + K7 5-band blend.

  python scripts/bench_canvas.py [--cols 25 --rows 20]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_canvas.py ...

With N ranks the canvas is split into N horizontal bands (uavm_canvas_set_band, 128-row recomputed halo, no exchange during
compute) and the finished bands are gathered on rank 0 with one NCCL gather.  Frames are synthesised on the device (one
texture, rolled per tile).  Prints one JSON line: per-stage device time (max over ranks), Mpx/s of finished canvas and the
compulsory-traffic rate (sum of W*H*3 over tiles + canvas*3 bytes) against the measured HBM peak."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.distributed as dist
import bench
from imagemosaicing_b200 import api, synth, dist as D, _lib as L

ap = argparse.ArgumentParser()
ap.add_argument("--cols", type=int, default=25); ap.add_argument("--rows", type=int, default=20)
ap.add_argument("--bands", type=int, default=5); ap.add_argument("--no-gather", action="store_true")
args = ap.parse_args()
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
W, H = bench.W, bench.H
n = args.cols * args.rows
rng = np.random.default_rng(20160308 + 5)
T = np.zeros((n, 9), np.float32)
for r in range(args.rows):
    for c in range(args.cols):
        k = r * args.cols + c
        a = 0.0 if k == 0 else np.deg2rad(rng.uniform(-2, 2)); s = 1.0 if k == 0 else rng.uniform(0.98, 1.02)
        T[k] = [s * np.cos(a), -s * np.sin(a), c * 1500.0, s * np.sin(a), s * np.cos(a), r * 1947.0, 0, 0, 1]
keep = np.ones(n, np.int32)
L.lib().uavm_resample_by_overlap(T.ctypes.data_as(L.f32p), n, W, H, api.C.c_float(0.7), keep.ctypes.data_as(L.i32p))   # ResampleByOverlap (0.7)
ctx = api.Context(local, torch.cuda.current_stream())
lay, chips = api.canvas_layout(T, keep, W, H)
cw, ch = lay.canvas_w, lay.canvas_h
chip_px = sum(chips[k].chip_w * chips[k].chip_h for k in range(n) if chips[k].keep)
need = n * W * H * 4 + chip_px * (4 + 1 + 4) + cw * ch * (8 / 3 + 4) + 2e9      # sources, chips + masks + chip pyramids, final canvas levels + result
free, total = torch.cuda.mem_get_info()
if need > 0.92 * free:
    raise SystemExit(f"needs ~{need / 1e9:.0f} GB, {free / 1e9:.0f} GB free: use fewer tiles (--cols/--rows)")
cv = api.Canvas(ctx, T, W, H, keep)
bands = D.canvas_bands(ch, world)
y0, y1 = bands[rank]
if world > 1:
    cv.set_band(y0, y1)
base = torch.from_numpy(synth.texture_image(rng, W, H, 6)).to(dev)
n_active = 0
for k in range(n):
    if keep[k] and cv.is_active(k):
        cv.set_image(k, torch.roll(base, shifts=(37 * k) % H, dims=0).contiguous()); n_active += 1
torch.cuda.synchronize()

def run():
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    e[0].record(); cv.warp(); e[1].record(); cv.seam_masks(); e[2].record(); cv.blend(args.bands); e[3].record()
    torch.cuda.synchronize()
    return [e[i].elapsed_time(e[i + 1]) for i in range(3)]
run()                                            # warm-up (allocates the canvas pyramid)
if world > 1: dist.barrier()
t = run()
gather_ms = 0.0
if world > 1 and not args.no_gather:
    max_rows = max(b[1] - b[0] for b in bands)
    send = torch.zeros((max_rows, cw, 3), dtype=torch.uint8, device=dev)
    cv.copy_result_rows(y0, y1, send)
    gathered = [torch.zeros_like(send) for _ in range(world)] if rank == 0 else None
    wsend = torch.zeros(1024, dtype=torch.uint8, device=dev)        # warm-up: NCCL sets its peer-to-peer channels up lazily
    dist.gather(wsend, [torch.zeros_like(wsend) for _ in range(world)] if rank == 0 else None, dst=0)
    torch.cuda.synchronize(); dist.barrier()
    g0 = torch.cuda.Event(enable_timing=True); g1 = torch.cuda.Event(enable_timing=True)
    g0.record(); dist.gather(send, gathered, dst=0); g1.record(); torch.cuda.synchronize()
    gather_ms = g0.elapsed_time(g1)
tt = torch.tensor(t + [gather_ms], device=dev, dtype=torch.float64)
if world > 1: dist.all_reduce(tt, op=dist.ReduceOp.MAX)
if rank == 0:
    warp_ms, seam_ms, blend_ms, gather_ms = [float(x) for x in tt]
    total_ms = warp_ms + seam_ms + blend_ms + gather_ms
    compulsory = int(keep.sum()) * W * H * 3 + cw * ch * 3
    model = chip_px * 40 + cw * ch * 32        # SURVEY §8d multi-pass model: ~40 B per fed chip pixel + ~32 B per canvas pixel
    pk = bench.peaks()
    print(json.dumps({"workload": f"configs[4]: {int(keep.sum())} warped tiles of {W}x{H} -> {cw}x{ch} canvas, {args.bands} bands", "n_gpus": world,
                      "warp_ms": warp_ms, "seam_masks_ms": seam_ms, "blend_ms": blend_ms, "gather_ms": gather_ms, "total_ms": total_ms,
                      "canvas_mpx_per_s": cw * ch / 1e6 / (total_ms / 1e3), "fed_chip_mpx": chip_px / 1e6,
                      "compulsory_gb": compulsory / 1e9, "compulsory_ms_at_hbm_peak": compulsory / 1e9 / (pk["hbm_gbs"] * world) * 1e3,
                      "multipass_model_gb": model / 1e9, "multipass_model_ms_at_hbm_peak": model / 1e9 / (pk["hbm_gbs"] * world) * 1e3,
                      "frac_of_multipass_model_roofline": model / 1e9 / (pk["hbm_gbs"] * world) * 1e3 / total_ms, "hbm_peak_gbs": pk["hbm_gbs"],
                      "active_tiles_rank0": n_active, "hbm_used_gb_rank0": (total - torch.cuda.mem_get_info()[0]) / 1e9}))
if world > 1: dist.destroy_process_group()
