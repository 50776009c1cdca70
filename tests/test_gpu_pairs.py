"""GPU parity tests (through the C ABI) of K2 match, K3 select and K4 RANSAC against the CPU oracle.

Bar (BASELINE.json north_star): match indices, candidate lists and RANSAC inlier masks bit-exact;
homography parameters within 1e-4 relative (they are in fact compared bit for bit first).
"""
import numpy as np
import pytest

from imagemosaicing_b200 import api, synth

pytestmark = pytest.mark.gpu


def _rand_desc(rng, n):
    return synth.sift_like_descriptors(rng, n)


@pytest.mark.parametrize("na,nb", [(1, 1), (5, 3), (127, 129), (128, 128), (256, 255), (257, 1000), (300, 2048), (2048, 2048), (1000, 4097)])
def test_match_parity_sizes(ctx, oracle, na, nb):
    rng = np.random.default_rng(na * 100003 + nb)
    A = _rand_desc(rng, na); B = _rand_desc(rng, nb)
    m = api.match(ctx, A.astype(np.float32), B.astype(np.float32))
    idx, d2 = oracle.match_l2(A, B)
    assert np.array_equal(m["queryIdx"], np.arange(na))
    assert np.array_equal(m["trainIdx"], idx)
    assert np.array_equal(m["distance"].view(np.uint32), np.sqrt(d2.astype(np.float32)).view(np.uint32))


def test_match_ties_duplicates_and_extremes(ctx, oracle):
    rng = np.random.default_rng(7)
    B = _rand_desc(rng, 700)
    B[100] = B[5]; B[650] = B[5]; B[300] = 0; B[301] = 0; B[400] = 255; B[401] = 255     # duplicates, zero rows, max rows
    A = np.concatenate([B[[5, 100, 300, 400, 699]], np.zeros((1, 128), np.uint8), np.full((1, 128), 255, np.uint8), _rand_desc(rng, 50)])
    fs = api.FeatureSet(ctx, [len(A), len(B)])
    fs.upload(0, A); fs.upload(1, B)
    pb = api.PairBatch(ctx, fs, [[0, 1]])
    pb.match()
    m = pb.matches(0)
    idx, d2 = oracle.match_l2(A, B)
    assert np.array_equal(m["trainIdx"], idx)          # lowest index wins ties: 5 (not 100/650), 300, 400
    assert m["trainIdx"][0] == 5 and m["trainIdx"][1] == 5 and m["trainIdx"][2] == 300 and m["trainIdx"][3] == 400
    assert np.array_equal((m["distance"].astype(np.float64) ** 2).round().astype(np.int64), d2.astype(np.int64))


def test_match_many_small_pairs_odd_tile_counts(ctx, oracle):
    """More work items than SMs, train images of 1, 2 and 3 tiles of 128 rows: every persistent CTA walks several items and the
    accumulator-buffer parity at the start of an item alternates (the two epilogue warp groups are bound to one TMEM buffer each)."""
    rng = np.random.default_rng(23)
    sizes = [100, 300, 129, 257, 384, 64, 200, 333]
    descs = [_rand_desc(rng, n) for n in sizes]
    descs[3][5] = descs[3][200]; descs[1][7] = descs[1][290]                       # duplicates across tiles: lowest index must win
    fs = api.FeatureSet(ctx, sizes)
    for i, d in enumerate(descs):
        fs.upload(i, d)
    pairs = [[i, j] for r in range(5) for i in range(len(sizes)) for j in range(len(sizes))]      # 320 pairs, ~450 work items
    pb = api.PairBatch(ctx, fs, pairs)
    pb.match()
    ref = {}
    for p, (i, j) in enumerate(pairs):
        if (i, j) not in ref:
            ref[(i, j)] = oracle.match_l2(descs[i], descs[j])
        idx, d2 = ref[(i, j)]
        m = pb.matches(p)
        assert np.array_equal(m["trainIdx"], idx), (p, i, j)
        assert np.array_equal(np.rint(m["distance"].astype(np.float64) ** 2).astype(np.int64), d2.astype(np.int64)), (p, i, j)


def test_match_batched_pairs_full_size(ctx, oracle):
    """BASELINE config sizes: 8192 keypoints per image, several pairs in one launch; checked against the
    oracle on two pairs and by a size-independent property on the rest (d2 recomputed from the indices)."""
    descs, kps, _ = synth.make_strip(5, 4000, 3000, 8192, seed=11)
    fs = api.FeatureSet(ctx, [len(d) for d in descs])
    for i, (d, k) in enumerate(zip(descs, kps)):
        fs.upload(i, d, k)
    pairs = [[0, 1], [1, 2], [2, 3], [3, 4], [4, 0], [2, 2]]
    pb = api.PairBatch(ctx, fs, pairs)
    pb.match()
    for p, (i, j) in enumerate(pairs):
        m = pb.matches(p)
        A = descs[i].astype(np.int64); B = descs[j].astype(np.int64)
        d2 = ((A - B[m["trainIdx"]]) ** 2).sum(1)
        assert np.array_equal(np.rint(m["distance"].astype(np.float64) ** 2).astype(np.int64), d2)
        if i == j:
            assert np.array_equal(m["trainIdx"], np.arange(len(A))) or (d2 == 0).all()
        if p < 2:
            idx, od2 = oracle.match_l2(descs[i], descs[j])
            assert np.array_equal(m["trainIdx"], idx)
            assert np.array_equal(d2, od2.astype(np.int64))


def test_match_32k_microbench_size(ctx, oracle):
    """BASELINE configs[3]: 32768 x 32768 SIFT-128 distance matrix (274.9 GFLOP, never materialised).  Checked by
    size-independent properties (distances recomputed from the returned indices; A against itself returns the
    identity) and against the oracle on a sample of query rows (full train set, so the arg-min is the global one)."""
    rng = np.random.default_rng(32768)
    n = 32768
    A = synth.sift_like_descriptors(rng, n); B = synth.sift_like_descriptors(rng, n)
    B[12345] = A[777]; B[23456] = A[777]                       # tie across far-apart tiles: lowest index wins
    fs = api.FeatureSet(ctx, [n, n])
    fs.upload(0, A, None); fs.upload(1, B, None)
    pb = api.PairBatch(ctx, fs, [[0, 1], [0, 0]])
    pb.match()
    m = pb.matches(0)
    d2 = ((A.astype(np.int64) - B[m["trainIdx"]].astype(np.int64)) ** 2).sum(1)
    assert np.array_equal(np.rint(m["distance"].astype(np.float64) ** 2).astype(np.int64), d2)
    assert m["trainIdx"][777] == 12345 and d2[777] == 0
    rows = rng.choice(n, 96, replace=False); rows[0] = 777
    idx, od2 = oracle.match_l2_fast(A[rows], B)
    assert np.array_equal(m["trainIdx"][rows], idx) and np.array_equal(d2[rows], od2.astype(np.int64))
    ms = pb.matches(1)                                          # A against itself
    ds = ((A.astype(np.int64) - A[ms["trainIdx"]].astype(np.int64)) ** 2).sum(1)
    assert (ds == 0).all() and (ms["trainIdx"] <= np.arange(n)).all()


@pytest.mark.parametrize("n,w,h", [(2048, 512, 512), (8192, 4000, 3000), (2000, 1000, 750), (30, 512, 512), (1335, 4000, 3000)])
def test_select_parity(ctx, oracle, n, w, h):
    rng = np.random.default_rng(n + w)
    kp1 = synth.random_keypoints(rng, n, w, h); kp2 = synth.random_keypoints(rng, n + 17, w, h)
    kp1[:5, 0] = w - 0.25          # exercises the nX == gridX quirk column (M/MosaicWithoutPos.cpp:4993-5010)
    train = rng.integers(0, n + 17, n).astype(np.int32)
    d2 = rng.integers(0, 60000, n).astype(np.int32)
    d2[rng.integers(0, n, n // 4)] = 12345           # many equal distances: (d2, queryIdx) order must hold
    m = np.zeros(n, dtype=np.dtype([("queryIdx", "<i4"), ("trainIdx", "<i4"), ("imgIdx", "<i4"), ("distance", "<f4")]))
    m["queryIdx"] = np.arange(n); m["trainIdx"] = train; m["distance"] = np.sqrt(d2.astype(np.float32))
    p1, p2 = api.select_match_pairs(ctx, m, kp1, kp2, w, h)
    o1, oi1, o2, oi2 = oracle.select(train, d2, kp1, kp2, w, h)
    assert len(p1) == len(oi1)
    assert np.array_equal(p1["id"], oi1) and np.array_equal(p2["id"], oi2)
    assert np.array_equal(np.c_[p1["x"], p1["y"]], o1) and np.array_equal(np.c_[p2["x"], p2["y"]], o2)


def _check_ransac(ctx, oracle, xy1, xy2, seed, sample_times=1000, dist=2.5):
    ok, mask, H, res = api.ransac2d(ctx, xy1, xy2, dist, sample_times, seed)
    o_ok, o_mask, o_H, o_n, st = oracle.ransac2d(xy1, xy2, dist, sample_times, seed)
    assert ok == o_ok
    assert res.n_inliers == o_n
    assert np.array_equal(mask, o_mask)
    assert res.max_support == st.max_support
    assert res.best_tuple == st.best_tuple
    assert res.n_tuples == st.n_tuples
    assert res.n_counted == st.n_counted
    if o_n >= 4:
        assert np.allclose(H, o_H, rtol=1e-4, atol=0), (H, o_H)       # north_star tolerance
        assert np.array_equal(H.view(np.uint32), o_H.view(np.uint32)) or np.array_equal(H, o_H), (H, o_H)
    return st


def test_ransac_parity_many_pairs(ctx, oracle):
    rng = np.random.default_rng(42)
    n_early = 0
    for trial in range(60):
        w, h = [(4000, 3000), (1000, 750), (512, 512)][trial % 3]
        n = int(rng.integers(4, 397))
        inl = 1.0 if trial % 10 == 9 else float(rng.uniform(0.15, 1.0))      # every 10th: all inliers -> early exit
        noise = float(rng.choice([0.0, 0.3, 0.5]))
        xy1, xy2, _ = synth.make_candidates(rng, n, w, h, inl, noise)
        st = _check_ransac(ctx, oracle, xy1, xy2, int(rng.integers(0, 2 ** 32)))
        n_early += st.early_exit
    assert n_early > 0          # the 0.99 early exit was exercised


def test_ransac_edge_cases(ctx, oracle):
    rng = np.random.default_rng(5)
    for n in (0, 1, 3, 4, 5, 8):
        xy1, xy2, _ = synth.make_candidates(rng, max(n, 1), 1000, 750, 1.0, 0.0)
        _check_ransac(ctx, oracle, xy1[:n], xy2[:n], 99)
    # pure noise: (almost) no consensus, loop runs until 5000 draws or 1000 counted
    xy1 = synth.random_keypoints(rng, 200, 4000, 3000); xy2 = synth.random_keypoints(rng, 200, 4000, 3000)
    _check_ransac(ctx, oracle, xy1, xy2, 1234)
    # all points collinear / duplicated: singular systems, exercises the slow path and the gates
    xy2 = np.stack([np.linspace(0, 999, 100), np.linspace(0, 500, 100)], 1).astype(np.float32)
    _check_ransac(ctx, oracle, xy2 + 3.0, xy2, 77)
    xy2 = np.repeat(synth.random_keypoints(rng, 10, 1000, 750), 10, 0)
    _check_ransac(ctx, oracle, xy2 * 1.01, xy2, 78)
    # small sample_times and the 5000 clamp
    xy1, xy2, _ = synth.make_candidates(rng, 300, 4000, 3000, 0.5, 0.5)
    _check_ransac(ctx, oracle, xy1, xy2, 5, sample_times=10)
    _check_ransac(ctx, oracle, xy1, xy2, 6, sample_times=7000)


def test_pipeline_match_select_ransac_strip(ctx, oracle):
    """End to end on a synthetic strip (C2-shaped, fewer images): every stage's output equals the oracle's."""
    w, h, nk = 4000, 3000, 8192
    descs, kps, Hs = synth.make_strip(4, w, h, nk, seed=20160308)
    fs = api.FeatureSet(ctx, [nk] * 4)
    for i in range(4):
        fs.upload(i, descs[i], kps[i])
    pairs = [[0, 1], [1, 2], [2, 3]]
    pb = api.PairBatch(ctx, fs, pairs)
    pb.match(); pb.select(w, h); pb.ransac(2.5, 1000, base_seed=1000)
    for p, (i, j) in enumerate(pairs):
        idx, d2 = oracle.match_l2(descs[i], descs[j])
        m = pb.matches(p)
        assert np.array_equal(m["trainIdx"], idx)
        o1, oi1, o2, oi2 = oracle.select(idx, d2, kps[i], kps[j], w, h)
        c1, c2 = pb.candidates(p)
        assert np.array_equal(c1["id"], oi1) and np.array_equal(c2["id"], oi2)
        o_ok, o_mask, o_H, o_n, st = oracle.ransac2d(o1, o2, 2.5, 1000, 1000 + p)
        mask, res = pb.ransac_result(p)
        assert res.n_inliers == o_n and np.array_equal(mask[:len(o_mask)], o_mask)
        assert res.best_tuple == st.best_tuple and res.max_support == st.max_support
        Hg = np.array(list(res.H), np.float32)
        assert np.allclose(Hg, o_H, rtol=1e-4, atol=0)
        # the recovered homography maps image j -> image i like the ground truth
        assert o_n > 100
        Ht = Hs[i] / Hs[i][2, 2]
        assert np.allclose(Hg[:8], Ht.reshape(-1)[:8], rtol=0.05, atol=2.0)
    out, n, acc = pb.collect(30)
    assert acc == 3 and n == sum(pb.ransac_result(p)[1].n_inliers for p in range(3))


def test_ransac_degenerate_inputs_generic_path(ctx, oracle):
    """Inputs that push many 4-tuples off the fast path (singular / ill-conditioned normal equations, skipped
    multipliers, missing pivots): the warp-cooperative generic evaluation must still match the oracle bit for bit."""
    rng = np.random.default_rng(17)
    cases = []
    # integer grid with duplicates and exact collinearities
    g = np.stack(np.meshgrid(np.arange(0, 1000, 125), np.arange(0, 750, 125)), -1).reshape(-1, 2).astype(np.float32)
    g = np.concatenate([g, g[:20]])
    cases.append((g * 1.0 + 2.0, g))
    # three collinear clusters
    t = np.linspace(0, 1, 60, dtype=np.float32)[:, None]
    c = np.concatenate([t * [900, 10] + [5, 5], t * [10, 700] + [5, 5], t * [900, 700] + [50, 20]]).astype(np.float32)
    cases.append((c + rng.normal(0, 0.2, c.shape).astype(np.float32), c))
    # everything inside a 3 px window (tiny baselines)
    s = rng.uniform(100, 103, (80, 2)).astype(np.float32)
    cases.append((s + 0.5, s))
    # very large coordinates (float32 normal equations overflow towards 1e20+)
    b = rng.uniform(0, 2e5, (120, 2)).astype(np.float32)
    cases.append((b * 0.999 + 30.0, b))
    # axis-aligned points: many exact zeros in the systems
    z = np.zeros((100, 2), np.float32); z[:50, 0] = rng.uniform(0, 4000, 50); z[50:, 1] = rng.uniform(0, 3000, 50)
    cases.append((z + 1.0, z))
    for k, (xy1, xy2) in enumerate(cases):
        _check_ransac(ctx, oracle, np.ascontiguousarray(xy1, np.float32), np.ascontiguousarray(xy2, np.float32), 4242 + k)


def test_pair_with_empty_image_yields_no_candidates(ctx):
    """An image without keypoints (allowed by uavm_mosaic_images) makes pairs with no matches: no candidates, RANSAC fails
    cleanly, the pair is rejected — in either role (query or train)."""
    rng = np.random.default_rng(8)
    nk = [300, 0, 400]
    descs = [rng.integers(0, 200, (n, 128)).astype(np.uint8) for n in nk]
    kps = [synth.random_keypoints(rng, n, 1000, 750) if n else np.zeros((0, 2), np.float32) for n in nk]
    fs = api.FeatureSet(ctx, nk)
    for i in range(3):
        if nk[i]:
            fs.upload(i, descs[i], kps[i])
    pb = api.PairBatch(ctx, fs, [[0, 1], [1, 2], [0, 2]])
    pb.match(); pb.select(1000, 750); pb.ransac(2.5, 100, base_seed=1)
    for p in (0, 1):
        c1, c2 = pb.candidates(p)
        assert len(c1) == 0 and len(c2) == 0
        mask, res = pb.ransac_result(p)
        assert res.n_inliers == 0 and res.ok == 0
    c1, _ = pb.candidates(2)
    assert len(c1) > 0
    out, n, acc = pb.collect(30)
    assert acc <= 1
