"""Multi-GPU canvas sharding check (run under torchrun on >= 2 GPUs; not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/multi_gpu_canvas_check.py

Every rank warps / masks / blends only its horizontal band of the mosaic canvas (+128-row halo, no exchange during
compute), the finished bands are gathered on rank 0 with one NCCL gather over NVLink, and rank 0 compares the
assembled mosaic byte for byte with the untiled single-GPU blend.  Also gathers the sharded pair results."""
import os
import sys
import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from imagemosaicing_b200 import api, synth, dist as D   # noqa: E402


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    ctx = api.Context(local, torch.cuda.current_stream())
    w, h, n, nk = 640, 480, 12, 2048
    descs, kps, Hs = synth.make_strip(n, w, h, nk, seed=5)
    T = [np.eye(3)]
    for Hk in Hs:
        A = Hk / Hk[2, 2]; A[2, :2] = 0
        T.append(T[-1] @ A)
    H = np.stack(T).astype(np.float32).reshape(n, 9)
    rng = np.random.default_rng(9)
    imgs = [synth.texture_image(rng, w, h, 5) for _ in range(n)]

    # ---- pair stage: pairs sharded round-robin, results gathered (NCCL all-reduce of disjoint slots) ----
    pairs = D.reference_pair_list(n, window=3)
    mine = D.shard_pairs(len(pairs), rank, world)
    fs = api.FeatureSet(ctx, [nk] * n)
    for i in range(n):
        fs.upload(i, descs[i], kps[i])
    pb = api.PairBatch(ctx, fs, pairs[mine])
    pb.match(); pb.select(w, h); pb.ransac(2.5, 1000, seeds=(1000 + mine).astype(np.uint32))
    recs = []
    for lp in range(len(mine)):
        mask, res = pb.ransac_result(lp)
        c1, c2 = pb.candidates(lp)
        keep = np.nonzero(mask[:len(c1)])[0] if res.n_inliers > 30 else np.zeros(0, np.int64)
        r = np.zeros(len(keep), D.MPP_DTYPE)
        r["xa"] = c1["x"][keep]; r["ya"] = c1["y"][keep]; r["ida"] = c1["id"][keep]; r["ia"] = pairs[mine[lp], 0]
        r["xb"] = c2["x"][keep]; r["yb"] = c2["y"][keep]; r["idb"] = c2["id"][keep]; r["ib"] = pairs[mine[lp], 1]
        recs.append(r)
    all_pairs = D.gather_match_pairs(recs, mine, len(pairs), device=dev)

    # ---- canvas stage: one band per rank ----
    cv = api.Canvas(ctx, H, w, h)
    cw, ch = cv.layout.canvas_w, cv.layout.canvas_h
    bands = D.canvas_bands(ch, world)
    y0, y1 = bands[rank]
    cv.set_band(y0, y1, 128)
    n_active = 0
    for k in range(n):
        if cv.is_active(k):
            cv.set_image(k, imgs[k]); n_active += 1
    cv.warp(); cv.seam_masks(); cv.blend(5)
    max_rows = max(b[1] - b[0] for b in bands)
    send = torch.zeros((max_rows, cw, 3), dtype=torch.uint8, device=dev)
    cv.copy_result_rows(y0, y1, send)
    torch.cuda.synchronize()
    gathered = [torch.zeros_like(send) for _ in range(world)] if rank == 0 else None
    dist.gather(send, gathered, dst=0)
    ok = 1
    if rank == 0:
        mosaic = np.zeros((ch, cw, 3), np.uint8)
        for r, (a, b) in enumerate(bands):
            mosaic[a:b] = gathered[r][:b - a].cpu().numpy()
        ref = api.Canvas(ctx, H, w, h)
        for k in range(n):
            ref.set_image(k, imgs[k])
        ref.warp(); ref.seam_masks(); ref.blend(5)
        full, _ = ref.result()
        same = np.array_equal(mosaic, full)
        # single-process pair results for comparison
        pb1 = api.PairBatch(ctx, fs, pairs)
        pb1.match(); pb1.select(w, h); pb1.ransac(2.5, 1000, seeds=(1000 + np.arange(len(pairs))).astype(np.uint32))
        out, n_mp, acc = pb1.collect(30)
        serial = np.frombuffer(out, dtype=D.MPP_DTYPE)[:n_mp]
        same_pairs = len(serial) == len(all_pairs) and np.array_equal(serial.view(np.uint8), all_pairs.view(np.uint8))
        print(f"multi_gpu_canvas_check: world={world} canvas={cw}x{ch} bands={bands} mosaic_equal={same} "
              f"pairs_equal={same_pairs} match_pairs={len(all_pairs)} accepted_pairs={acc}")
        ok = 1 if (same and same_pairs) else 0
    flag = torch.tensor([ok], device=dev)
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
