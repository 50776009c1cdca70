"""End-to-end check on the reference's own sample run (SURVEY §4, e2e tier): SIFT features of Release/test_data/DSC00004..23
(tests/golden/ref_images_sift.npz, extracted with cv2 and the reference's SIFT parameters by tests/golden/make_sift_fixture.py)
go through the GPU pair path (match -> select -> RANSAC over the reference's candidate rule), the accept rule, connectivity
and the global affine alignment; the result is compared with what the reference's run left behind: feature_temp/matchPairs.match
(accepted pairs) and tran0.txt (transforms).  SIFT build, matcher (exact vs FLANN) and RNG differ from the author's run, so the
gate is the survey's: linear terms within 0.03, translations within 25 px."""
import ctypes as C
import os
import numpy as np
import pytest

from imagemosaicing_b200 import api, _lib as L, dist as D

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _features():
    g = np.load(os.path.join(GOLD, "ref_images_sift.npz"))
    cnt = g["counts"]; off = np.concatenate([[0], np.cumsum(cnt)])
    kp = [g["kp"][off[i]:off[i + 1]] for i in range(len(cnt))]
    de = [g["desc"][off[i]:off[i + 1]] for i in range(len(cnt))]
    return kp, de, int(g["width"]), int(g["height"])


def test_reference_sample_run(ctx, oracle):
    kp, de, w, h = _features()
    n = len(kp)
    assert n == 20 and (w, h) == (1000, 750)
    fs = api.FeatureSet(ctx, [len(d) for d in de])
    for i in range(n):
        fs.upload(i, de[i], kp[i])
    pairs = D.reference_pair_list(n)                       # j in (i, min(n, i + 182)): all 190 pairs
    assert len(pairs) == 190
    pb = api.PairBatch(ctx, fs, pairs)
    pb.match(); pb.select(w, h); pb.ransac(2.5, 1000, base_seed=20160308)
    # parity with the CPU oracle on a few pairs of real descriptors
    for p in (0, 57, 189):
        i, j = pairs[p]
        idx, d2 = oracle.match_l2_fast(de[i], de[j])
        assert np.array_equal(pb.matches(p)["trainIdx"], idx)
        x1, i1, x2, i2 = oracle.select(idx, d2, kp[i], kp[j], w, h)
        c1, c2 = pb.candidates(p)
        assert np.array_equal(c1["id"], i1) and np.array_equal(c2["id"], i2)
        ok, mask, Hm, ni, st = oracle.ransac2d(x1, x2, 2.5, 1000, (20160308 + p) & 0xffffffff)
        gm, res = pb.ransac_result(p)
        assert res.n_inliers == ni and np.array_equal(gm[:len(mask)], mask)
    out, n_m, n_acc = pb.collect(30)
    mp = (L.MatchPointPairs * n_m).from_buffer_copy(bytes(out)[:n_m * 40])
    ours = set((r.ptA_i, r.ptB_i) for r in mp)
    ref_pairs = set((r.ptA_i, r.ptB_i) for r in api.read_match_file(os.path.join(GOLD, "ref_matchPairs.match")))
    assert len(ref_pairs) == 58 and n_acc == len(ours)
    assert len(ours) >= 38 and len(ours & ref_pairs) >= 0.9 * len(ours)      # the accepted graph is (almost) a subgraph of the author's
    # connectivity + reference image 0 fixed + global affine alignment (host side of the library)
    label = (C.c_int32 * n)()
    assert L.lib().uavm_connected_images(mp, n_m, n, label) == 0
    assert list(label) == [1] * n
    for r in mp:
        if r.ptA_i == 0: r.ptA_Fixed = 1
        if r.ptB_i == 0: r.ptB_Fixed = 1
    init = (L.ImageTransform * n)(); res_t = (L.ImageTransform * n)()
    for i in range(n):
        for t in range(9):
            init[i].h.m[t] = 1.0 if t in (0, 4, 8) else 0.0
        init[i].fixed = 1 if i == 0 else 0
    assert L.lib().uavm_align_affine(mp, n_m, init, n, 1, res_t) == 0
    T = np.array([[res_t[i].h.m[t] for t in range(9)] for i in range(n)], np.float32)
    ref_T, _ = api.read_transform_file(os.path.join(GOLD, "ref_tran0.txt"))
    d = T[1:, :6] - ref_T[1:, :6]
    lin = np.abs(d[:, [0, 1, 3, 4]]).max(); tr = np.abs(d[:, [2, 5]])
    print(f"reference sample run: {len(ours)} accepted pairs ({len(ours & ref_pairs)} of them among the author's 58), {n_m} inlier matches, "
          f"max |linear diff| {lin:.4f}, translation diff median {np.median(tr):.2f} px max {tr.max():.2f} px")
    assert lin <= 0.03 and tr.max() <= 25.0
