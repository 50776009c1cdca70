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
