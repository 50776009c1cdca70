"""Multi-GPU check through the C ABI (run under torchrun on >= 2 GPUs; tests/test_gpu_multi.py launches it when the box has
them):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/multi_gpu_canvas_check.py

Pairs are sharded round-robin and merged with uavm_pairbatch_allgather (device-side pack, ncclAllGather, device-side
compaction); every rank warps / masks / blends only its rectangle of the mosaic canvas (row bands, then a 2-D grid), the
finished rectangles are moved to rank 0 with uavm_canvas_gather (peer copies into rank 0's mosaic mapped with CUDA IPC, or written
there by the blend itself after uavm_canvas_bind_root), and rank 0 compares the match list and the assembled mosaics byte for byte with the single-GPU results."""
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
    nd = api.Dist.from_torch(ctx)
    out_all, n_all, acc_all = nd.allgather_matches(pb, len(pairs), 30)
    all_pairs = np.frombuffer(out_all, dtype=D.MPP_DTYPE)[:n_all]

    # ---- canvas stage: one rectangle per rank (row bands, then a 2-D grid), gathered on rank 0 through the C ABI ----
    # third pass: the 2-D grid again with the root bound (uavm_canvas_bind_root) — the blend's level-0 kernel writes the
    # rectangles into rank 0's mosaic over NVLink itself and the gather is only the completion barrier; blended twice
    mosaics = []
    for grid2d, bound in ((False, False), (True, False), (True, True)):
        cv = api.Canvas(ctx, H, w, h)
        cw, ch = cv.layout.canvas_w, cv.layout.canvas_h
        rects = D.canvas_grid(cw, ch, world) if grid2d else [(0, a, cw, b) for (a, b) in D.canvas_bands(ch, world)]
        cv.set_rect(*rects[rank])
        for k in range(n):
            if cv.is_active(k):
                cv.set_image(k, imgs[k])
        if bound:
            nd.bind_canvas_root(cv, root=0)
        cv.warp(); cv.seam_masks()
        for _ in range(2 if bound else 1):
            cv.blend(5)
            nd.gather_canvas(cv, rects, root=0)
        ctx.sync()
        mosaics.append(cv.result()[0] if rank == 0 else None)
        dist.barrier()                                  # rank 0 has read its mosaic before anybody unmaps / frees
        cv.close()
    ok = 1
    if rank == 0:
        ref = api.Canvas(ctx, H, w, h)
        for k in range(n):
            ref.set_image(k, imgs[k])
        ref.warp(); ref.seam_masks(); ref.blend(5)
        full, _ = ref.result()
        same = all(np.array_equal(m, full) for m in mosaics)
        # single-process pair results for comparison
        pb1 = api.PairBatch(ctx, fs, pairs)
        pb1.match(); pb1.select(w, h); pb1.ransac(2.5, 1000, seeds=(1000 + np.arange(len(pairs))).astype(np.uint32))
        out, n_mp, acc = pb1.collect(30)
        serial = np.frombuffer(out, dtype=D.MPP_DTYPE)[:n_mp]
        same_pairs = len(serial) == len(all_pairs) and np.array_equal(serial.view(np.uint8), all_pairs.view(np.uint8))
        same_pairs = same_pairs and acc == acc_all
        print(f"multi_gpu_canvas_check: world={world} canvas={cw}x{ch} mosaic_equal={same} "
              f"pairs_equal={same_pairs} match_pairs={len(all_pairs)} accepted_pairs={acc}")
        ok = 1 if (same and same_pairs) else 0
    flag = torch.tensor([ok], device=dev)
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
