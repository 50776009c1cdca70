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


def test_f32_weight_pyrdown_close_and_exact_on_binary_masks(oracle):
    rng = np.random.default_rng(1)
    f = rng.random((64, 96)).astype(np.float32)
    assert np.abs(oracle.pyr_down_f32(f) - cv2.pyrDown(f)).max() <= 2.4e-7     # last-ulp (SIMD order in OpenCV)
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
    d = np.abs(out.astype(np.int32) - r8.astype(np.int32))
    assert np.array_equal(om, rm)
    assert d.max() <= 1
    assert (d > 0).mean() < 0.02        # measured ~0.8 %: ulp differences of the f32 weight pyramid at levels 4-5
    assert om.mean() > 100              # most of the canvas is covered
