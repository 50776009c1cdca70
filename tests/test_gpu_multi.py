"""Multi-GPU behind the C ABI (csrc/dist.cu).  The single-rank cases run on any GPU box; the torchrun case needs >= 2 GPUs and
is skipped otherwise (gpurun --gpus 2 runs it)."""
import os
import subprocess
import sys
import numpy as np
import pytest

from imagemosaicing_b200 import api, synth, dist as D

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_allgather_single_rank_equals_collect(ctx):
    """uavm_pairbatch_allgather with world == 1 (device-side pack -> NCCL all-gather -> device-side compaction) returns exactly
    uavm_pairbatch_collect's list; a rejected pair (too few inliers) and the counts-only form are covered."""
    w, h, nk, n = 1000, 750, 2048, 6
    descs, kps, Hs = synth.make_strip(n, w, h, nk, seed=77)
    fs = api.FeatureSet(ctx, [nk] * n)
    for i in range(n):
        fs.upload(i, descs[i], kps[i])
    pairs = D.reference_pair_list(n, window=4)            # includes non-overlapping pairs that RANSAC rejects
    pb = api.PairBatch(ctx, fs, pairs)
    pb.match(); pb.select(w, h)
    ctx.fork(); pb.ransac(2.5, 1000, base_seed=5); ctx.unfork()       # results fetched without an explicit join
    nd = api.Dist(ctx, 0, 1, api.Dist.unique_id())
    out_a, n_a, acc_a = nd.allgather_matches(pb, len(pairs), 30)
    out_c, n_c, acc_c = pb.collect(30)
    assert (n_a, acc_a) == (n_c, acc_c) and 0 < acc_c < len(pairs)
    a = np.frombuffer(out_a, dtype=D.MPP_DTYPE)[:n_a]; c = np.frombuffer(out_c, dtype=D.MPP_DTYPE)[:n_c]
    assert np.array_equal(a.view(np.uint8), c.view(np.uint8))
    nd.close()


def test_two_contexts_in_one_process(ctx):
    """Two uavm_ctx in one process (the C++ host pattern: one context per worker thread): kernels with opted-in shared
    memory run from both, and from a context created after the first one already launched."""
    import torch
    ctx2 = api.Context(1 if torch.cuda.device_count() > 1 else 0)           # another device when the box has one
    rng = np.random.default_rng(3)
    a = rng.integers(0, 200, (300, 128)).astype(np.float32); b = rng.integers(0, 200, (500, 128)).astype(np.float32)
    m1 = api.match(ctx, a, b); m2 = api.match(ctx2, a, b)
    assert np.array_equal(m1["trainIdx"], m2["trainIdx"])
    xy1, xy2, _ = synth.make_candidates(rng, 200, 1000, 750, 0.5, 0.5)
    r1 = api.ransac2d(ctx, xy1, xy2, 2.5, 1000, 9); r2 = api.ransac2d(ctx2, xy1, xy2, 2.5, 1000, 9)
    assert np.array_equal(r1[1], r2[1])
    ctx2.close()


def test_multi_gpu_check_under_torchrun():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    world = 4 if n >= 4 else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29613", os.path.join(ROOT, "tests", "multi_gpu_canvas_check.py")]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:]
    assert "mosaic_equal=True" in r.stdout and "pairs_equal=True" in r.stdout
