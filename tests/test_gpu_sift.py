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


@pytest.mark.parametrize("w,h,seed", [(1000, 750, 1), (640, 480, 2)])
def test_sift_vs_cv2(ctx, w, h, seed):
    rng = np.random.default_rng(seed)
    img = synth.texture_image(rng, w, h, 7)
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
    assert frac >= 0.97
    assert dd.mean() < 0.25 and np.percentile(dd.max(axis=1), 95) <= 4
    # deterministic
    kp_g2, d_g2 = s.detect_and_compute(img)
    assert np.array_equal(kp_g2.view(np.uint8), kp_g.view(np.uint8)) and np.array_equal(d_g2, d_g)
    s.close()
