"""End-to-end GPU test of the MosaicVavImages-shaped shim (uavm_mosaic_images) against the same pipeline
composed from the CPU oracle stages: every stage is bit-exact, so transforms and mosaic bytes must be equal."""
import numpy as np
import pytest

from imagemosaicing_b200 import api, synth

pytestmark = pytest.mark.gpu


def _scene(n, w, h, nk, seed):
    descs, kps, Hs = synth.make_strip(n, w, h, nk, seed=seed)
    rng = np.random.default_rng(seed + 5)
    base = synth.texture_image(rng, w, h, 6)
    images = [np.ascontiguousarray(np.roll(base, 13 * k, axis=1)) for k in range(n)]
    return images, descs, kps, Hs


@pytest.mark.parametrize("blending", [2, 0])
def test_mosaic_images_equals_oracle_pipeline(ctx, oracle, blending):
    if oracle.ref() is None:
        pytest.skip("needs oracle/_ref (reference overlap filter)")
    from oracle import pipeline as OP
    n, w, h, nk = 5, 640, 480, 2048
    images, descs, kps, Hs = _scene(n, w, h, nk, 77)
    out, T, fixed = api.mosaic_images(ctx, images, [d.astype(np.float32) for d in descs], kps,
                                      {"blending": blending, "pairWindow": 3, "seed": 123}, 1.0)
    o_T, o_out = OP.mosaic_images(images, descs, kps, pair_window=3, seed=123, blending=blending)
    assert np.array_equal(T, o_T)
    assert out.shape == o_out.shape
    assert np.array_equal(out, o_out), f"{(out != o_out).sum()} differing bytes"
    assert fixed[0] == 1 and (fixed[1:] == 0).all()
    # the chain of ground-truth pair homographies is recovered to a few pixels
    G = np.eye(3)
    for k in range(1, n):
        G = G @ (Hs[k - 1] / Hs[k - 1][2, 2])
        assert np.allclose(T[k, :6].reshape(2, 3)[:, :2], G[:2, :2], atol=0.02)
        assert np.allclose(T[k, [2, 5]], G[:2, 2], atol=6.0)


def test_mosaic_images_argument_errors(ctx):
    n, w, h, nk = 3, 320, 240, 256
    images, descs, kps, _ = _scene(n, w, h, nk, 5)
    d32 = [d.astype(np.float32) for d in descs]
    with pytest.raises(api.UavmError):                      # fewer than two images: -1 (M/MosaicWithoutPos.cpp:10157-10165)
        api.mosaic_images(ctx, images[:1], d32[:1], kps[:1])
    with pytest.raises(api.UavmError):                      # unrelated images: no pair accepted: -2
        rng = np.random.default_rng(0)
        bad = [synth.sift_like_descriptors(rng, nk).astype(np.float32) for _ in range(n)]
        api.mosaic_images(ctx, images, bad, kps)


def test_load_match_pairs_resume_path(ctx, tmp_path):
    """loadMatchPairs = 1 (M/MosaicWithoutPos.cpp:4465-4477): the matches of a first run, written to and read back from a
    matchPairs.match file, reproduce the same transforms and the same mosaic without the matching stage."""
    n, w, h, nk = 5, 640, 480, 2048
    images, descs, kps, _ = _scene(n, w, h, nk, 78)
    prm = {"blending": 2, "pairWindow": 3, "seed": 9}
    out, T, fixed, matches = api.mosaic_images(ctx, images, [d.astype(np.float32) for d in descs], kps, prm, 1.0, return_matches=True)
    assert matches is not None and len(matches) > 100
    path = str(tmp_path / "matchPairs.match")
    api.write_match_file(path, matches)
    loaded = api.read_match_file(path)
    assert bytes(loaded) == bytes(matches)
    out2, T2, fixed2 = api.mosaic_from_matches(ctx, images, loaded, prm, 1.0)
    assert np.array_equal(T, T2) and np.array_equal(fixed, fixed2) and np.array_equal(out, out2)
    # transforms survive tran0.txt to its 6 printed digits
    tpath = str(tmp_path / "tran0.txt")
    api.write_transform_file(tpath, T, fixed)
    T3, _ = api.read_transform_file(tpath)
    assert np.allclose(T3[:, :6], T[:, :6], rtol=2e-5, atol=1e-6)
    with pytest.raises(api.UavmError):                      # image index out of range in the loaded list: -1
        bad = type(loaded).from_buffer_copy(loaded); bad[0].ptA_i = 99
        api.mosaic_from_matches(ctx, images, bad, prm, 1.0)


def test_block_alignment_recovers_ground_truth(ctx, oracle):
    """configs[2] in small: a 3 x 4 UAV block with a shared world-point model (pairs along and across strips, loops in the
    pair graph).  All overlapping pairs go through the GPU pair path, the accepted matches through connectivity and the
    global affine alignment; the recovered transforms must reproduce the ground-truth poses to about a pixel, and the
    match list must be the one the CPU oracle produces (same pairs, same inliers, same order)."""
    import ctypes as C
    from imagemosaicing_b200 import _lib as L
    rows, cols, w, h, nk = 3, 4, 640, 480, 2048
    descs, kps, poses, pairs = synth.make_block(rows, cols, w, h, nk, seed=11)
    n = rows * cols
    fs = api.FeatureSet(ctx, [nk] * n)
    for i in range(n):
        fs.upload(i, descs[i], kps[i])
    pb = api.PairBatch(ctx, fs, pairs)
    pb.match(); pb.select(w, h); pb.ransac(2.5, 1000, seeds=(1000 + np.arange(len(pairs))).astype(np.uint32))
    out, n_m, n_acc = pb.collect(30)
    mp = (L.MatchPointPairs * n_m).from_buffer_copy(bytes(out)[:n_m * 40])
    # oracle composition of the same pair path
    o_rows = []
    for p, (i, j) in enumerate(pairs):
        idx, d2 = oracle.match_l2_fast(descs[i], descs[j])
        x1, i1, x2, i2 = oracle.select(idx, d2, kps[i], kps[j], w, h)
        ok, mask, Hm, ni, st = oracle.ransac2d(x1, x2, 2.5, 1000, 1000 + p)
        if ni > 30:
            for k in np.nonzero(mask)[0]:
                o_rows.append((i, x1[k, 0], x1[k, 1], j, x2[k, 0], x2[k, 1]))
    assert n_m == len(o_rows)
    for r, o in zip(mp, o_rows):
        assert (r.ptA_i, r.ptA.x, r.ptA.y, r.ptB_i, r.ptB.x, r.ptB.y) == o
    label = (C.c_int32 * n)()
    assert L.lib().uavm_connected_images(mp, n_m, n, label) == 0 and list(label) == [1] * n
    for r in mp:
        if r.ptA_i == 0: r.ptA_Fixed = 1
        if r.ptB_i == 0: r.ptB_Fixed = 1
    init = (L.ImageTransform * n)(); res = (L.ImageTransform * n)()
    for i in range(n):
        for t in range(9):
            init[i].h.m[t] = 1.0 if t in (0, 4, 8) else 0.0
        init[i].fixed = 1 if i == 0 else 0
    assert L.lib().uavm_align_affine(mp, n_m, init, n, 1, res) == 0
    corners = np.array([[0, 0], [w - 1, 0], [w - 1, h - 1], [0, h - 1]], np.float64)
    err = []
    for k in range(n):
        G = np.linalg.inv(poses[0]) @ poses[k]
        T = np.array([res[k].h.m[t] for t in range(9)], np.float64).reshape(3, 3)
        err.append(np.abs(synth.apply_h(G, corners) - synth.apply_h(T, corners)).max())
    assert max(err) < 2.0 and np.median(err) < 1.0, (max(err), np.median(err))


def test_mosaic_sequence_chunks_like_mosaic_uav_video(ctx):
    """uavm_mosaic_sequence = the chunk loop of MosaicUavVideo (M/MosaicWithoutPos.cpp:10252-10300): every chunk of
    max_once frames equals uavm_mosaic_images on that slice; chunks start at n + numMosaiced."""
    n, w, h, nk = 7, 480, 360, 1536
    images, descs, kps, _ = _scene(n, w, h, nk, 31)
    d32 = [d.astype(np.float32) for d in descs]
    prm = {"blending": 2, "pairWindow": 3, "seed": 5}
    chunks = api.mosaic_sequence(ctx, images, d32, kps, 3, prm, 1.0)
    assert [c[0] for c in chunks] == [0, 3, 6] or [c[0] for c in chunks] == [0, 3]      # the last chunk holds one frame: no mosaic
    for first, img in chunks:
        cnt = min(3, n - first)
        if cnt < 2:
            assert img is None
            continue
        ref, _, _ = api.mosaic_images(ctx, images[first:first + cnt], d32[first:first + cnt], kps[first:first + cnt], prm, 1.0)
        assert img is not None and np.array_equal(img, ref)
