"""CPU tests of the K7 oracle (oracle/oracle_blend.c) against cv2 4.13 — the available proxy for the
reference's OpenCV 2.4.0 MultiBandBlender (third-party binary, not under /root/reference)."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")


def test_int16_pyramid_primitives_bit_exact(oracle):
    rng = np.random.default_rng(0)
    for (h, w) in [(64, 96), (33, 47), (32, 32), (2, 2), (1, 8)]:
        a = rng.integers(-3000, 3000, (h, w, 3)).astype(np.int16)
        assert np.array_equal(oracle.pyr_down_s16(a), cv2.pyrDown(a))
        assert np.array_equal(oracle.pyr_up_s16(a), cv2.pyrUp(a))
    # extremes saturate like OpenCV
    a = np.full((16, 16, 3), 32767, np.int16); a[::2] = -32768
    assert np.array_equal(oracle.pyr_down_s16(a), cv2.pyrDown(a))
    assert np.array_equal(oracle.pyr_up_s16(a), cv2.pyrUp(a))


def test_f32_weight_pyrdown_bit_exact(oracle):
    """The operation order of OpenCV's f32 pyrDown (vector row / column forms + scalar borders and tails) is part of the
    result; the oracle reproduces cv2.pyrDown bit for bit on random f32 data at every width class."""
    rng = np.random.default_rng(1)
    shapes = [(64, 96), (33, 47), (40, 70), (16, 18), (1, 1), (2, 3), (5, 4), (7, 9), (9, 10), (3, 21)]
    shapes += [(int(rng.integers(1, 40)), int(rng.integers(1, 60))) for _ in range(40)]
    for (h, w) in shapes:
        f = rng.random((h, w)).astype(np.float32)
        assert np.array_equal(oracle.pyr_down_f32(f), cv2.pyrDown(f).reshape((h + 1) // 2, (w + 1) // 2)), (h, w)
        g = (rng.random((h, w)).astype(np.float32) * np.float32(1e-3))           # small magnitudes: same order, other exponents
        assert np.array_equal(oracle.pyr_down_f32(g), cv2.pyrDown(g).reshape((h + 1) // 2, (w + 1) // 2)), (h, w)
    f = rng.random((64, 96)).astype(np.float32)
    b = (rng.random((64, 96)) > 0.5).astype(np.float32)
    l1 = oracle.pyr_down_f32(b); l2 = oracle.pyr_down_f32(l1); l3 = oracle.pyr_down_f32(l2)
    c1 = cv2.pyrDown(b); c2 = cv2.pyrDown(c1); c3 = cv2.pyrDown(c2)
    assert np.array_equal(l1, c1) and np.array_equal(l2, c2) and np.array_equal(l3, c3)   # dyadic: exact through level 3


def test_blend_oracle_vs_cv2_blender(oracle):
    from imagemosaicing_b200 import synth
    rng = np.random.default_rng(2)
    w, h, n = 320, 240, 3
    T = [np.eye(3)]
    for k in range(1, n):
        Hk = synth.pair_homography(rng, w, h, overlap=(0.55, 0.8)); Hk[2, :2] = 0
        T.append(T[-1] @ Hk)
    H = np.stack(T).astype(np.float32).reshape(n, 9)
    imgs = [synth.texture_image(rng, w, h, 5) for _ in range(n)]
    canvas, chips_l = oracle.canvas_layout(H, None, w, h)
    chips, masks = [], []
    for k in range(n):
        px, m = oracle.warp_chip(imgs[k], canvas, chips_l[k]); chips.append(px); masks.append(m)
    seam = oracle.seam_masks(masks, [chips_l[k] for k in range(n)], canvas.canvas_w, canvas.canvas_h)
    tls = [(chips_l[k].beg_x, chips_l[k].beg_y) for k in range(n)]
    out, om = oracle.multiband_blend(chips, seam, tls, canvas.canvas_w, canvas.canvas_h, 5)
    b = cv2.detail_MultiBandBlender(0, 5)
    b.prepare((0, 0, canvas.canvas_w, canvas.canvas_h))
    for k in range(n):
        b.feed(chips[k].astype(np.int16), seam[k], tls[k])
    rs, rm = b.blend(None, None)
    r8 = np.clip(rs, 0, 255).astype(np.uint8)
    assert np.array_equal(om, rm)
    assert np.array_equal(out, r8)      # bit exact: the f32 weight pyrDown follows OpenCV's operation order
    assert om.mean() > 100              # most of the canvas is covered


def test_blend_oracle_bit_exact_random_masks_and_bands(oracle):
    """Random chips, random placements, binary AND grey (non-dyadic weight) masks, 1..5 bands: the oracle equals
    cv2.detail_MultiBandBlender byte for byte."""
    rng = np.random.default_rng(5)
    for trial in range(6):
        cw, ch = int(rng.integers(150, 500)), int(rng.integers(120, 400))
        n = 4; nb = int(rng.integers(1, 6))
        chips, masks, tls = [], [], []
        for k in range(n):
            w = int(rng.integers(40, cw // 2 + 40)); h = int(rng.integers(40, ch // 2 + 40))
            chips.append(rng.integers(0, 256, (h, w, 3)).astype(np.uint8))
            masks.append(rng.integers(0, 256, (h, w)).astype(np.uint8) if trial % 2 else (rng.random((h, w)) > 0.3).astype(np.uint8) * 255)
            tls.append((int(rng.integers(0, cw - w + 1)), int(rng.integers(0, ch - h + 1))))
        out, om = oracle.multiband_blend(chips, masks, tls, cw, ch, nb)
        b = cv2.detail_MultiBandBlender(0, nb)
        b.prepare((0, 0, cw, ch))
        for k in range(n):
            b.feed(chips[k].astype(np.int16), masks[k], tls[k])
        rs, rm = b.blend(None, None)
        assert np.array_equal(om, rm), trial
        assert np.array_equal(out, np.clip(rs, 0, 255).astype(np.uint8)), trial


def test_normalise_shortcut_is_exact():
    """K7 replaces short(d / (wsum + 1e-5f)) by d - sign(d) when wsum == 1.0f (and by 0 when d == 0): identical for every
    int16 d in f32 arithmetic (csrc/blend.cu: norm_div)."""
    d = np.arange(-32768, 32768, dtype=np.int32)
    w = np.float32(1.0) + np.float32(1e-5)
    v = np.trunc(d.astype(np.float32) / w).astype(np.int32)
    assert np.array_equal(v, d - np.sign(d))
    for ws in (np.float32(0.0), np.float32(0.37), np.float32(2.5)):
        assert np.trunc(np.float32(0.0) / (ws + np.float32(1e-5))) == 0
