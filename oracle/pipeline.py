"""CPU composition of the oracle stages in the order of CMosaicByPose::MosaicWithoutPose
(M/MosaicWithoutPos.cpp:4430-4679) -> LaplacianPyramidBlending (M/MosaicImage.cpp:2205-2510).
TEST INFRASTRUCTURE ONLY (the checker for uavm_mosaic_images)."""
import ctypes as C
import numpy as np
from . import oracle as O


def mosaic_images(images, descs, kps, ransac_dist=2.5, sample_times=1000, pair_window=182, min_inner=30,
                  seed=20160308, scale=1.0, num_bands=5, blending=2, overlap_t=0.7):
    n = len(images); h, w = images[0].shape[:2]
    pairs = [(i, j) for i in range(n) for j in range(i + 1, min(n, i + pair_window))]
    matches = []          # rows: imgA xA yA fixedA imgB xB yB fixedB
    for p, (i, j) in enumerate(pairs):
        idx, d2 = O.match_l2(descs[i], descs[j])
        x1, i1, x2, i2 = O.select(idx, d2, kps[i], kps[j], w, h)
        ok, mask, H, ninl, st = O.ransac2d(x1, x2, ransac_dist, sample_times, (seed + p) & 0xffffffff)
        if ninl > min_inner:
            for k in np.nonzero(mask)[0]:
                matches.append([i, x1[k, 0], x1[k, 1], 0, j, x2[k, 0], x2[k, 1], 0])
    if not matches:
        return None
    m = np.array(matches, np.float64)
    # largest connected component (the tests use graphs without size ties)
    parent = list(range(n))
    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]; a = parent[a]
        return a
    for r in m:
        parent[find(int(r[0]))] = find(int(r[4]))
    comps = {}
    for i in range(n):
        comps.setdefault(find(i), []).append(i)
    touched = set(int(v) for v in m[:, 0]) | set(int(v) for v in m[:, 4])
    best = max((c for c in comps.values() if any(x in touched for x in c)), key=len)
    label = np.zeros(n, np.int32); label[best] = 1
    keep_rows = []
    rows = [list(r) for r in m]
    k = 0
    while k < len(rows):                                   # same swap-with-last removal as the reference
        if label[int(rows[k][0])] == 0 or label[int(rows[k][4])] == 0:
            rows[k] = rows[-1]; rows.pop()
        else:
            k += 1
    m = np.array(rows, np.float64)
    m[m[:, 0] == 0, 3] = 1; m[m[:, 4] == 0, 7] = 1
    fixed = (label == 0).astype(np.int32); fixed[0] = 1
    rc, T = O.align_affine(m, fixed)
    assert rc == 0
    T[label == 0, 8] = 0
    Hs = T.copy(); Hs[:, :6] *= np.float32(scale)
    if blending != 2:
        return T, O.paste(Hs, images)
    keep = np.ones(n, np.int32)
    f32p = C.POINTER(C.c_float); i32p = C.POINTER(C.c_int32)
    Hc = np.ascontiguousarray(Hs, np.float32)
    O.ref().ref_resample_by_overlap(Hc.ctypes.data_as(f32p), n, w, h, C.c_float(overlap_t), keep.ctypes.data_as(i32p))
    canvas, chips = O.canvas_layout(Hs, keep, w, h)
    cs, ms, valid = [], [], []
    for k in range(n):
        if chips[k].keep:
            px, msk = O.warp_chip(images[k], canvas, chips[k]); cs.append(px); ms.append(msk); valid.append(chips[k])
    seam = O.seam_masks(ms, valid, canvas.canvas_w, canvas.canvas_h)
    out, om = O.multiband_blend(cs, seam, [(c.beg_x, c.beg_y) for c in valid], canvas.canvas_w, canvas.canvas_h, num_bands)
    return T, out
