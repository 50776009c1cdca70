"""Multi-GPU plumbing for the pair stage: image pairs are independent units, so ranks take
disjoint slices of the pair list with NO data-path collective; one gather of the (small) per-pair inlier
records at the end gives every rank the full MatchPointPairs list, after which the global alignment runs
identically on every rank ("replicas only").  Works with torch.distributed backends nccl (GPU tensors) and
gloo (CPU tests)."""
import numpy as np

MPP_DTYPE = np.dtype([("xa", "<f4"), ("ya", "<f4"), ("ida", "<i4"), ("ia", "<i4"), ("fa", "<i4"),
                      ("xb", "<f4"), ("yb", "<f4"), ("idb", "<i4"), ("ib", "<i4"), ("fb", "<i4")])   # 40-byte MatchPointPairs


def reference_pair_list(n_images, window=182):
    """j in (i, min(n, i + window)) — the reference's candidate rule (M/MosaicWithoutPos.cpp:5083-5084)."""
    return np.array([(i, j) for i in range(n_images) for j in range(i + 1, min(n_images, i + window))], np.int32).reshape(-1, 2)


def shard_pairs(n_pairs, rank, world):
    """Round-robin shard: pair p belongs to rank p % world.  Returns the global pair indices of this rank."""
    return np.arange(rank, n_pairs, world, dtype=np.int64)


def gather_match_pairs(local_records, local_pair_ids, n_pairs, group=None, device=None):
    """local_records: list (one per local pair, aligned with local_pair_ids) of MPP_DTYPE arrays (possibly empty).
    Returns the concatenation over ALL pairs in global pair order — identical on every rank and identical to
    what a single process produces."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    counts = torch.zeros(n_pairs, dtype=torch.int64)
    for pid, rec in zip(local_pair_ids, local_records):
        counts[int(pid)] = len(rec)
    dev = device if device is not None else torch.device("cpu")
    counts = counts.to(dev)
    if world > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    counts = counts.cpu().numpy()
    offsets = np.concatenate([[0], np.cumsum(counts)])
    total = int(offsets[-1])
    buf = np.zeros(total, MPP_DTYPE)
    for pid, rec in zip(local_pair_ids, local_records):
        buf[offsets[int(pid)]:offsets[int(pid) + 1]] = rec
    if world > 1 and total > 0:
        # every slot is written by exactly one rank and zero elsewhere: a SUM over the raw int32 words merges them
        t = torch.from_numpy(buf.view(np.int32).copy()).to(dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        buf = t.cpu().numpy().view(MPP_DTYPE).copy()
    return buf


def split_collected(records, local_pairs):
    """uavm_pairbatch_collect returns the accepted pairs' inlier records concatenated in local pair order; this splits them
    back into one array per local pair (empty for rejected pairs), the form gather_match_pairs takes.
    records: MPP_DTYPE array; local_pairs: (n, 2) image indices of this rank's pairs, in the PairBatch's order."""
    out = []
    k = 0
    n = len(records)
    for i, j in local_pairs:
        k0 = k
        while k < n and records["ia"][k] == i and records["ib"][k] == j:
            k += 1
        out.append(records[k0:k])
    if k != n:
        raise ValueError("records are not grouped in local pair order")
    return out


def canvas_bands(canvas_h, world, align=32):
    """Split the canvas rows into `world` horizontal bands whose edges are multiples of `align`
    (the last band ends at canvas_h).  Returns [(y0, y1)] per rank; ranks beyond the available rows get (0, 0)."""
    n_units = (canvas_h + align - 1) // align
    out = []
    for r in range(world):
        u0 = (n_units * r) // world; u1 = (n_units * (r + 1)) // world
        y0 = u0 * align; y1 = min(u1 * align, canvas_h)
        out.append((y0, y1) if y1 > y0 else (0, 0))
    return out


def canvas_grid(canvas_w, canvas_h, world, align=32):
    """Split the canvas into `world` rectangles of a gx x gy grid (gx * gy == world, as square as the canvas allows:
    SURVEY §8e's 4 x 2 grid for 8 GPUs) whose inner edges are multiples of `align`.  Returns [(x0, y0, x1, y1)] per rank,
    row-major; a 2-D grid halves the chips a rank touches compared with full-width bands when chips are taller than a band."""
    best = None
    for gx in range(1, world + 1):
        if world % gx:
            continue
        gy = world // gx
        cost = abs(canvas_w / gx - canvas_h / gy)          # prefer square cells
        if best is None or cost < best[0]:
            best = (cost, gx, gy)
    _, gx, gy = best
    def cuts(n, g):
        units = (n + align - 1) // align
        return [min((units * k // g) * align, n) for k in range(g)] + [n]
    xs, ys = cuts(canvas_w, gx), cuts(canvas_h, gy)
    return [(xs[i], ys[j], xs[i + 1], ys[j + 1]) for j in range(gy) for i in range(gx)]


def canvas_grid_balanced(canvas_w, canvas_h, world, chip_boxes, align=32, base_cost=2.0):
    """Like canvas_grid, but the cuts equalise WORK instead of area: a canvas cell costs base_cost + the number of chips covering
    it (the seam masks evaluate every covering chip per pixel, the blend visits them, the warp and the pyramids follow what a
    rectangle owns).  Column cuts first, then every column group gets its own row cuts, so the rectangles still partition the
    canvas.  chip_boxes: iterable of (beg_x, beg_y, chip_w, chip_h) of the kept chips in canvas coordinates."""
    import numpy as np
    best = None
    for gx in range(1, world + 1):
        if world % gx:
            continue
        gy = world // gx
        c = abs(canvas_w / gx - canvas_h / gy)
        if best is None or c < best[0]:
            best = (c, gx, gy)
    _, gx, gy = best
    wu, hu = (canvas_w + align - 1) // align, (canvas_h + align - 1) // align
    cover = np.zeros((hu + 1, wu + 1), np.float64)                  # 2-D difference array of the chip boxes, in cells
    for (bx, by, w, h) in chip_boxes:
        x0, y0 = max(int(bx) // align, 0), max(int(by) // align, 0)
        x1, y1 = min((int(bx) + int(w) + align - 1) // align, wu), min((int(by) + int(h) + align - 1) // align, hu)
        if x1 <= x0 or y1 <= y0:
            continue
        cover[y0, x0] += 1; cover[y0, x1] -= 1; cover[y1, x0] -= 1; cover[y1, x1] += 1
    cost = cover.cumsum(0).cumsum(1)[:hu, :wu] + base_cost

    def cuts(weights, g, n):
        cum = np.concatenate([[0.0], np.cumsum(weights)])
        out = [0]
        for k in range(1, g):
            u = int(np.searchsorted(cum, cum[-1] * k / g))
            u = min(max(u, out[-1] // align + 1), len(weights) - (g - k))      # strictly increasing, room for the rest
            out.append(min(u * align, n))
        return out + [n]

    xs = cuts(cost.sum(0), gx, canvas_w)
    rects = [None] * world
    for i in range(gx):
        c0, c1 = xs[i] // align, (xs[i + 1] + align - 1) // align
        ys = cuts(cost[:, c0:c1].sum(1), gy, canvas_h)
        for j in range(gy):
            rects[j * gx + i] = (xs[i], ys[j], xs[i + 1], ys[j + 1])
    return rects
