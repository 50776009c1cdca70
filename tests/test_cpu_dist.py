"""world_size-2 gloo test of the multi-GPU host logic: sharding the pair list over ranks and gathering the
per-pair inlier records reproduces the single-process MatchPointPairs list (order and bytes)."""
import os
import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from imagemosaicing_b200 import dist as D


def _fake_pair_records(pid, i, j):
    rng = np.random.default_rng(1000 + pid)
    n = int(rng.integers(0, 60))
    rec = np.zeros(n, D.MPP_DTYPE)
    rec["xa"] = rng.uniform(0, 4000, n); rec["ya"] = rng.uniform(0, 3000, n); rec["ida"] = rng.integers(0, 8192, n)
    rec["xb"] = rng.uniform(0, 4000, n); rec["yb"] = rng.uniform(0, 3000, n); rec["idb"] = rng.integers(0, 8192, n)
    rec["ia"] = i; rec["ib"] = j
    return rec


def _worker(rank, world, port, pairs, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = D.shard_pairs(len(pairs), rank, world)
    recs = [_fake_pair_records(int(p), *pairs[int(p)]) for p in mine]
    full = D.gather_match_pairs(recs, mine, len(pairs))
    np.save(os.path.join(out_dir, f"full_{rank}.npy"), full)
    dist.destroy_process_group()


def test_pair_sharding_and_gather_world2(tmp_path):
    pairs = D.reference_pair_list(12, window=4)
    assert len(pairs) == sum(min(12, i + 4) - i - 1 for i in range(12))
    world = 2
    shards = [D.shard_pairs(len(pairs), r, world) for r in range(world)]
    assert sorted(np.concatenate(shards).tolist()) == list(range(len(pairs)))      # disjoint cover
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, pairs, str(tmp_path)), nprocs=world, join=True)
    serial = np.concatenate([_fake_pair_records(p, *pairs[p]) for p in range(len(pairs))])
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), f"full_{r}.npy"))
        assert got.dtype == D.MPP_DTYPE and got.itemsize == 40
        assert np.array_equal(got.view(np.uint8), serial.view(np.uint8))


def test_split_collected_roundtrip():
    """collect() concatenates the accepted pairs' records in local pair order; split_collected undoes it (rejected pairs
    become empty arrays), including consecutive pairs that share an image."""
    pairs = np.array([(0, 1), (0, 2), (1, 2), (2, 5), (5, 6)], np.int32)
    per_pair = [_fake_pair_records(p, *pairs[p]) for p in range(len(pairs))]
    per_pair[1] = per_pair[1][:0]                                     # a rejected pair
    cat = np.concatenate(per_pair)
    back = D.split_collected(cat, pairs)
    assert len(back) == len(pairs)
    for a, b in zip(back, per_pair):
        assert np.array_equal(a.view(np.uint8), b.view(np.uint8))
    try:
        D.split_collected(cat[::-1].copy(), pairs)
        assert False, "out-of-order records must be rejected"
    except ValueError:
        pass


def test_reference_pair_window():
    p = D.reference_pair_list(200, 182)
    assert len(p) == 19729                       # the reference rule on a 200-image block
    assert len(D.reference_pair_list(50, 2)) == 49


def test_canvas_grid_balanced_partitions_the_canvas():
    """The work-balanced grid is still a partition of the canvas into `world` rectangles with aligned inner edges."""
    import numpy as np
    from imagemosaicing_b200 import dist as D
    rng = np.random.default_rng(4)
    cw, ch = 10007, 7013
    boxes = [(int(rng.integers(0, cw - 900)), int(rng.integers(0, ch - 700)), 900, 700) for _ in range(60)]
    for world in (1, 2, 3, 4, 6, 8):
        rects = D.canvas_grid_balanced(cw, ch, world, boxes)
        assert len(rects) == world
        cover = np.zeros((ch // 7 + 1, cw // 7 + 1), np.int32)       # sampled coverage count
        area = 0
        for (x0, y0, x1, y1) in rects:
            assert 0 <= x0 < x1 <= cw and 0 <= y0 < y1 <= ch
            assert x0 % 32 == 0 and y0 % 32 == 0 and (x1 % 32 == 0 or x1 == cw) and (y1 % 32 == 0 or y1 == ch)
            area += (x1 - x0) * (y1 - y0)
            ys = np.arange(0, ch, 7); xs = np.arange(0, cw, 7)
            cover[np.ix_((ys >= y0) & (ys < y1), (xs >= x0) & (xs < x1))] += 1
        assert area == cw * ch and cover[:len(np.arange(0, ch, 7)), :len(np.arange(0, cw, 7))].min() == 1 and cover.max() == 1
