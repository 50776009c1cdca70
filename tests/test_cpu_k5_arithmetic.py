"""The two exact-arithmetic identities the K5 warp kernels rest on (csrc/canvas.cu), checked on the CPU with numpy so that
their validity does not hinge on the sampled images of the GPU parity tests:

  1. u8 -> f32 conversion + first multiply as ONE fused multiply-add:  fl32(fma(2^23 + g, w, -(2^23 w))) == fl32(float(g) * w)
     for every byte g and every float weight w in [0, 1] (the bilinear weights p, 1-p).  (2^23 + g) w and 2^23 w are both
     exact in binary64, their difference is exactly g w, so the single rounding of the FMA is the rounding of the product.
  2. int(v) for 0 <= v < 2^23 as add.rz(v, 2^23): the sum rounded toward zero is 2^23 + floor(v), so the integer is the low
     mantissa bits and (sum - 2^23) == float(int(v)).
"""
import numpy as np


def _weights(rng, n):
    p = rng.random(n).astype(np.float32)                                       # fractions of source coordinates
    edge = np.array([0.0, 1.0, np.float32(1) - np.float32(2) ** -24, np.float32(2) ** -24, 0.5, 0.25, np.float32(1e-7)], np.float32)
    coarse = (rng.integers(0, 4096, n // 4) / np.float32(4096)).astype(np.float32)   # what ys - int(ys) looks like for ys ~ 3000
    return np.concatenate([p, np.float32(1.0) - p, edge, coarse])


def test_fma_on_two23_plus_byte_is_the_rounded_product():
    rng = np.random.default_rng(0)
    w = _weights(rng, 200000)
    for g in range(256):
        ref = (np.float32(g) * w).astype(np.float32)                            # what the reference computes: float(g) * w, rounded once
        G = np.float64(8388608.0 + g)                                           # 2^23 + g, exact
        exact = G * w.astype(np.float64) - np.float64(8388608.0) * w.astype(np.float64)   # each term exact in binary64 (<= 48 significant bits)
        assert np.array_equal(exact, np.float64(g) * w.astype(np.float64))     # the difference is exactly g * w ...
        assert np.array_equal(exact.astype(np.float32), ref)                   # ... so one rounding to binary32 gives the reference's product


def _add_rz_two23(v):
    """binary32 add.rz(v, 2^23) for 0 <= v < 2^23, emulated exactly: the exact sum has < 53 significant bits, so binary64 holds
    it; rounding toward zero to the binary32 grid of [2^23, 2^24) (spacing 1) is floor()."""
    s = v.astype(np.float64) + 8388608.0
    return np.floor(s).astype(np.float32)


def test_add_rz_two23_is_truncation():
    rng = np.random.default_rng(1)
    v = np.concatenate([rng.random(100000).astype(np.float32) * 4000, rng.random(100000).astype(np.float32) * 256,
                        np.array([0.0, 0.99999994, 1.0, 254.99998, 255.0, 255.99998, 2999.9998, 8388607.5], np.float32),
                        np.arange(0, 4000, dtype=np.float32), np.nextafter(np.arange(1, 4000, dtype=np.float32), np.float32(0))])
    t = _add_rz_two23(v)
    bits = t.view(np.uint32)
    assert np.array_equal((bits - np.uint32(0x4B000000)).astype(np.int64), v.astype(np.int64))        # low mantissa bits = int(v) (C truncation)
    assert np.array_equal(t - np.float32(8388608.0), np.trunc(v))                                     # (sum - 2^23) == float(int(v)), exact
    assert np.array_equal(bits & np.uint32(0xff), (v.astype(np.int64) & 0xff).astype(np.uint32))      # the output byte is the low byte


def test_coordinate_chain_is_monotone():
    """The footprint of a chip tile is derived from its 4 corner pixels: xs(x, y) = fl(fl(fl(xt * a) + fl(yt * b)) + c) with
    xt = fl(fl(fl(x) - dgx) - sx + bx) must be monotone in x and in y for fixed signs of a, b (every rounding is monotone)."""
    rng = np.random.default_rng(2)
    for _ in range(200):
        a, b, c = [np.float32(v) for v in rng.uniform(-1.3, 1.3, 3) * np.array([1, 1, 3000])]
        dgx, sx, bx = np.float32(rng.uniform(0, 2000)), np.float32(rng.uniform(-1, 0)), np.float32(rng.integers(0, 5000))
        x = np.arange(0, 257, dtype=np.float32)
        xt = ((x - dgx) - sx) + bx
        for yt in (np.float32(-17.25), np.float32(1234.5)):
            xs = ((xt * a).astype(np.float32) + np.float32(yt * b)).astype(np.float32) + c
            d = np.diff(xs.astype(np.float64))
            assert (d >= 0).all() or (d <= 0).all()
