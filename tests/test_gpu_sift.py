"""f1: GPU SIFT (csrc/sift.cu) against cv2 4.13 SIFT_create(2000, 3, 0.01, 20) — the available proxy for the reference's OpenCV
2.4.0 nonfree SIFT (M/MosaicWithoutPos.cpp:4852-4872; third-party binary, parity with 2.4.0 itself is unpinned).
Keypoints are matched by position / scale / orientation; descriptors are compared on the matched keypoints."""
import numpy as np
import pytest

from imagemosaicing_b200 import api, synth

pytestmark = pytest.mark.gpu
cv2 = pytest.importorskip("cv2")


def _match(kp_g, kp_c):
    """For every cv2 keypoint the nearest GPU keypoint of the same octave/layer; returns (index or -1) per cv2 keypoint."""
    out = np.full(len(kp_c), -1, np.int64)
    gx = kp_g["x"]; gy = kp_g["y"]
    for i, k in enumerate(kp_c):
        d = np.hypot(gx - k.pt[0], gy - k.pt[1])
        cand = np.nonzero((d < 0.05) & ((kp_g["octave"] & 0xffff) == (k.octave & 0xffff)))[0]
        best, bd = -1, 1e9
        for j in cand:
            da = abs(kp_g["angle"][j] - k.angle); da = min(da, 360 - da)
            if da < 1.0 and d[j] + da < bd:
                best, bd = j, d[j] + da
        out[i] = best
    return out


def _scene(rng, w, h, dense):
    """dense: blurred noise on top of the texture -> several thousand extrema (exercises retainBest)."""
    img = synth.texture_image(rng, w, h, 7)
    if dense:
        n = cv2.GaussianBlur(rng.normal(0, 1, (h, w, 1)).astype(np.float32), (0, 0), 1.3)
        n = n[..., None] if n.ndim == 2 else n
        img = np.clip(img.astype(np.float32) * 0.6 + 50 + 260 * n, 0, 255).astype(np.uint8)
    return img


@pytest.mark.parametrize("w,h,seed,dense", [(1000, 750, 1, False), (640, 480, 2, False), (1000, 750, 3, True), (517, 389, 4, True)])
def test_sift_vs_cv2(ctx, w, h, seed, dense):
    rng = np.random.default_rng(seed)
    img = _scene(rng, w, h, dense)
    ref = cv2.SIFT_create(2000, 3, 0.01, 20)
    kp_c, d_c = ref.detectAndCompute(img, None)
    s = api.Sift(ctx, w, h, 2000, 3, 0.01, 20.0, 1.6)
    kp_g, d_g = s.detect_and_compute(img)
    assert abs(len(kp_g) - len(kp_c)) <= 0.02 * len(kp_c)
    assert np.array_equal(d_g, np.round(d_g)) and d_g.min() >= 0 and d_g.max() <= 255
    m = _match(kp_g, kp_c)
    frac = (m >= 0).mean()
    ok = m >= 0
    dd = np.abs(d_g[m[ok]] - d_c[ok])
    print(f"sift {w}x{h}: cv2 {len(kp_c)} gpu {len(kp_g)} matched {frac:.4f}; descriptor |diff| max {dd.max():.0f} mean {dd.mean():.4f} "
          f"exact rows {(dd.max(axis=1) == 0).mean():.3f}; pos err max {np.hypot(kp_g['x'][m[ok]] - np.array([k.pt[0] for k in kp_c])[ok], kp_g['y'][m[ok]] - np.array([k.pt[1] for k in kp_c])[ok]).max():.5f}")
    if dense:
        assert len(kp_c) >= 2000                      # retainBest was exercised
    assert frac >= 0.99
    assert dd.mean() < 0.05 and dd.max() <= 3 and (dd.max(axis=1) == 0).mean() > 0.9
    # deterministic
    kp_g2, d_g2 = s.detect_and_compute(img)
    assert np.array_equal(kp_g2.view(np.uint8), kp_g.view(np.uint8)) and np.array_equal(d_g2, d_g)
    s.close()


def test_mosaic_images_sift_end_to_end(ctx):
    """MosaicVavImages' true signature: images in, mosaic out.  Overlapping views of one scene (known shifts and small
    rotations) -> GPU SIFT -> match / RANSAC / global alignment / warp / blend; the recovered transforms map every view onto
    the first one within a pixel."""
    rng = np.random.default_rng(11)
    W0, H0, w, h, n = 1500, 700, 512, 384, 5
    scene = _scene(rng, W0, H0, True)
    views, gt = [], []
    for k in range(n):
        a = np.deg2rad(rng.uniform(-3, 3)) if k else 0.0
        tx, ty = 60 + 200 * k, 150 + (rng.uniform(-20, 20) if k else 0)
        M = np.array([[np.cos(a), -np.sin(a), tx], [np.sin(a), np.cos(a), ty]], np.float64)       # view pixel -> scene pixel
        views.append(cv2.warpAffine(scene, M, (w, h), flags=cv2.INTER_LINEAR | cv2.WARP_INVERSE_MAP, borderMode=cv2.BORDER_REFLECT))
        gt.append(np.vstack([M, [0, 0, 1]]))
    out, T, fixed, matches = api.mosaic_images_sift(ctx, views)
    assert matches is not None and len(matches) > 200
    assert out.shape[1] > w * 2 and out.mean() > 20
    corners = np.array([[0, 0], [w - 1, 0], [w - 1, h - 1], [0, h - 1]], np.float64)
    for k in range(1, n):
        G = np.linalg.inv(gt[0]) @ gt[k]                                   # view k -> view 0
        Tk = T[k].astype(np.float64).reshape(3, 3); Tk[2] = [0, 0, 1]
        err = np.abs(synth.apply_h(G, corners) - synth.apply_h(Tk, corners)).max()
        assert err < 1.5, (k, err)
