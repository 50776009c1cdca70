"""Host logic of K7: the region planner (imagemosaicing_b200/csrc/blend_plan.h) against a brute-force model of what the blend
reads.  The planner decides which pyramid values are computed (C_i) and where a chip is evaluated (U_i); a rectangle that is too
small would silently change the mosaic, so the closure properties are checked exhaustively on small random cases:
  * U_i contains every level-i pixel whose mask-pyramid weight can be non-zero inside the canvas rectangle S_i
    (pyrDown of the weights reads 2x-2..2x+2 with reflect-101),
  * C_i contains U_i, everything the pyrUp of the 4 x 2 blocks touching U_{i-1} reads ((x>>1)-1..(x>>1)+1, clamped),
    and everything the pyrDown that produces C_{i+1} reads,
  * S_{i+1} contains everything the pyrUp of the 4 x 2 blocks of S_i reads; S_0 contains the output rectangle.
MultiBandBlender::feed / blend (driven by M/MosaicImage.cpp:2476-2486) is what these rectangles shortcut."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    out = tmp_path_factory.mktemp("plan") / "libplan_shim.so"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-shared", "-fPIC", "-I", os.path.join(ROOT, "imagemosaicing_b200", "csrc"),
                           os.path.join(ROOT, "tests", "plan_shim", "plan_shim.cpp"), "-o", str(out)])
    return C.CDLL(str(out))


def _reflect101(p, n):
    if n == 1:
        return 0
    while p < 0 or p >= n:
        p = -p if p < 0 else 2 * n - 2 - p
    return p


def _down_taps(q, n):                     # level i pixels read by pyrDown output q (level size n)
    return {_reflect101(t, n) for t in range(2 * q - 2, 2 * q + 3)}


def _up_taps(x, n):                       # level i+1 pixels read by pyrUp output x (coarse level size n): reflect near, replicate far
    c = x >> 1
    lo = (1 if n > 1 else 0) if c == 0 else c - 1
    return {lo, min(c, n - 1), min(c + 1, n - 1)}


def _contains(rect, xs, ys):
    x0, y0, x1, y1 = rect
    return all(x0 <= x < x1 for x in xs) and all(y0 <= y < y1 for y in ys)


def _plan(shim, cw, ch, nb, out, chip, a0):
    pc = (C.c_int * (4 + 6 * (nb + 1)))()
    shim.shim_plan_canvas(cw, ch, nb, *out, pc)
    S = [tuple(pc[4 + 6 * i + 2:4 + 6 * i + 6]) for i in range(nb + 1)]
    lw = [pc[4 + 6 * i] for i in range(nb + 1)]; lh = [pc[4 + 6 * i + 1] for i in range(nb + 1)]
    po = (C.c_int * (7 + 10 * (nb + 1)))()
    shim.shim_plan_chip(cw, ch, nb, *out, *chip, *a0, po)
    roi = dict(active=po[0], tlx=po[1], tly=po[2], width=po[3], height=po[4], top=po[5], left=po[6])
    lv = [dict(pw=po[7 + 10 * i], ph=po[8 + 10 * i], U=tuple(po[9 + 10 * i:13 + 10 * i]), C=tuple(po[13 + 10 * i:17 + 10 * i])) for i in range(nb + 1)]
    return (pc[0], pc[1]), lw, lh, S, roi, lv


def test_canvas_level_rectangles_are_closed_under_pyrup(shim):
    rng = np.random.default_rng(1)
    for _ in range(60):
        nb = int(rng.integers(1, 5))
        cw, ch = int(rng.integers(40, 300)), int(rng.integers(40, 300))
        x0, y0 = 2 * int(rng.integers(0, cw // 4)), 2 * int(rng.integers(0, ch // 4))
        x1 = cw if rng.random() < 0.3 else min(cw, x0 + 2 * int(rng.integers(1, cw // 2)))
        y1 = ch if rng.random() < 0.3 else min(ch, y0 + 2 * int(rng.integers(1, ch // 2)))
        (W, H), lw, lh, S, _, _ = _plan(shim, cw, ch, nb, (x0, y0, x1, y1), (0, 0, 8, 8), (0, 0, 8, 8))
        assert W % (1 << nb) == 0 and H % (1 << nb) == 0 and W >= cw and H >= ch
        assert S[0][0] <= x0 and S[0][1] <= y0 and S[0][2] >= x1 and S[0][3] >= y1
        for i in range(nb):
            sx0, sy0, sx1, sy1 = S[i]
            assert 0 <= sx0 < sx1 <= lw[i] + 1 and 0 <= sy0 < sy1 <= lh[i] + 1
            # the level kernels work on 4 x 2 blocks aligned to 4 columns / 2 rows of the level
            bx0, bx1 = sx0 & ~3, (sx1 + 3) & ~3
            by0, by1 = sy0 & ~1, (sy1 + 1) & ~1
            xs = set().union(*[_up_taps(x, lw[i + 1]) for x in range(bx0, bx1)])
            ys = set().union(*[_up_taps(y, lh[i + 1]) for y in range(by0, by1)])
            assert _contains(S[i + 1], xs, ys), (i, S[i], S[i + 1])


def test_chip_rectangles_cover_what_the_blend_reads(shim):
    rng = np.random.default_rng(2)
    n_active = 0
    for it in range(150):
        nb = int(rng.integers(1, 5))
        cw, ch = int(rng.integers(60, 260)), int(rng.integers(60, 260))
        sharded = rng.random() < 0.5
        if sharded:
            x0, y0 = 2 * int(rng.integers(0, cw // 4)), 2 * int(rng.integers(0, ch // 4))
            x1, y1 = min(cw, x0 + 2 * int(rng.integers(4, cw // 2))), min(ch, y0 + 2 * int(rng.integers(4, ch // 2)))
        else:
            x0, y0, x1, y1 = 0, 0, cw, ch
        w, h = min(int(rng.integers(8, 90)), cw), min(int(rng.integers(8, 90)), ch)
        bx, by = int(rng.integers(0, cw - w + 1)), int(rng.integers(0, ch - h + 1))
        ax0, ay0 = int(rng.integers(0, w)), int(rng.integers(0, h))
        ax1, ay1 = int(rng.integers(ax0 + 1, w + 1)), int(rng.integers(ay0 + 1, h + 1))
        (W, H), lw, lh, S, roi, lv = _plan(shim, cw, ch, nb, (x0, y0, x1, y1), (bx, by, w, h), (ax0, ay0, ax1, ay1))
        assert roi["tlx"] % (1 << nb) == 0 and roi["tly"] % (1 << nb) == 0 and roi["width"] % (1 << nb) == 0 and roi["height"] % (1 << nb) == 0
        assert roi["tlx"] + roi["left"] == bx and roi["tly"] + roi["top"] == by
        assert roi["tlx"] + roi["width"] <= W and roi["tly"] + roi["height"] <= H
        # brute force: support of the weight pyramid (separable: per axis), level by level
        sx = set(range(roi["left"] + ax0, roi["left"] + ax1)); sy = set(range(roi["top"] + ay0, roi["top"] + ay1))
        any_contrib = False
        for i in range(nb + 1):
            pw, ph = lv[i]["pw"], lv[i]["ph"]
            if i > 0:
                sx = {q for q in range(pw) if _down_taps(q, lv[i - 1]["pw"]) & sx}
                sy = {q for q in range(ph) if _down_taps(q, lv[i - 1]["ph"]) & sy}
            ox, oy = roi["tlx"] >> i, roi["tly"] >> i
            cx = {x for x in sx if S[i][0] <= x + ox < S[i][2]}; cy = {y for y in sy if S[i][1] <= y + oy < S[i][3]}
            if cx and cy:
                any_contrib = True
                assert _contains(lv[i]["U"], cx, cy), ("U", it, i, lv[i]["U"], min(cx), max(cx), min(cy), max(cy))
        assert bool(roi["active"]) == any_contrib or roi["active"]        # the planner may be conservative, never the opposite
        if not roi["active"]:
            continue
        n_active += 1
        for i in range(nb + 1):
            U, Cc = lv[i]["U"], lv[i]["C"]
            pw, ph = lv[i]["pw"], lv[i]["ph"]
            if U[2] > U[0] and U[3] > U[1]:
                assert Cc[0] <= U[0] and Cc[1] <= U[1] and Cc[2] >= U[2] and Cc[3] >= U[3], ("U in C", i)
            if Cc[2] > Cc[0]:
                assert 0 <= Cc[0] and Cc[2] <= pw and 0 <= Cc[1] and Cc[3] <= ph
                if 1 <= i < nb:
                    assert Cc[0] % 2 == 0 and Cc[1] % 2 == 0 and Cc[2] % 2 == 0 and Cc[3] % 2 == 0
            if i < nb:
                Cn = lv[i + 1]["C"]
                # pyrUp of the 4 x 2 blocks (aligned in CANVAS level coordinates) that touch U_i
                if U[2] > U[0] and U[3] > U[1]:
                    ox, oy = roi["tlx"] >> i, roi["tly"] >> i
                    bx0, bx1 = ((U[0] + ox) & ~3) - ox, ((U[2] + ox + 3) & ~3) - ox
                    by0, by1 = ((U[1] + oy) & ~1) - oy, ((U[3] + oy + 1) & ~1) - oy
                    xs = set().union(*[_up_taps(x, lv[i + 1]["pw"]) for x in range(max(bx0, 0), min(bx1, pw))])
                    ys = set().union(*[_up_taps(y, lv[i + 1]["ph"]) for y in range(max(by0, 0), min(by1, ph))])
                    assert _contains(Cn, xs, ys), ("pyrUp", it, i, U, Cn)
                # pyrDown that produces C_{i+1} reads level i inside C_i (level 0: inside the ROI, the chip is read with its own border rule)
                if Cn[2] > Cn[0]:
                    xs = set().union(*[_down_taps(q, pw) for q in range(Cn[0], Cn[2])])
                    ys = set().union(*[_down_taps(q, ph) for q in range(Cn[1], Cn[3])])
                    assert _contains(Cc, xs, ys), ("pyrDown", it, i, Cc, Cn)
    assert n_active > 80
