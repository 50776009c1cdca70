"""CPU tests (`-m "not gpu"`): the oracle against the reference's golden vectors and against the reference's
own compiled RANSAC core, the host-compiled kernel math against the oracle, the host-side stages of the
product library, and the C-ABI surface.  No GPU compute is called here."""
import ctypes as C
import os
import re
import struct

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


# ------------------------------------------------------------------------------------------------
# golden fixtures of the reference's one shipped run
# ------------------------------------------------------------------------------------------------
def _load_match_txt():
    return np.loadtxt(os.path.join(GOLD, "ref_matchPairs.txt"))        # imgA xA yA fixedA imgB xB yB fixedB


def _load_tran0():
    return np.loadtxt(os.path.join(GOLD, "ref_tran0.txt"))             # rows for images 1..19: m0..m7 fixed


def test_golden_match_file_consistency():
    """matchPairs.match: int32 n + n x 40-byte MatchPointPairs; 5918 records / 58 accepted pairs, each with
    more than MIN_INNER_POINTS = 30 inliers (M/MosaicWithoutPos.cpp:5049,5201)."""
    raw = open(os.path.join(GOLD, "ref_matchPairs.match"), "rb").read()
    n = struct.unpack("<i", raw[:4])[0]
    assert n == 5918 and len(raw) == 4 + n * 40
    rec = np.frombuffer(raw[4:], dtype=np.dtype([("xa", "<f4"), ("ya", "<f4"), ("ida", "<i4"), ("ia", "<i4"), ("fa", "<i4"),
                                                 ("xb", "<f4"), ("yb", "<f4"), ("idb", "<i4"), ("ib", "<i4"), ("fb", "<i4")]))
    pairs = {}
    for r in rec:
        pairs.setdefault((int(r["ia"]), int(r["ib"])), []).append(r)
    assert len(pairs) == 58
    assert min(len(v) for v in pairs.values()) > 30
    txt = _load_match_txt()
    assert len(txt) == n
    # per-pair homography refit (double DLT): residuals consistent with the 2.5 px RANSAC gate
    worst = 0.0
    for (ia, ib), v in pairs.items():
        a = np.array([[r["xa"], r["ya"]] for r in v], np.float64); b = np.array([[r["xb"], r["yb"]] for r in v], np.float64)
        A = []
        for (x1, y1), (x2, y2) in zip(a, b):
            A.append([x2, y2, 1, 0, 0, 0, -x1 * x2, -x1 * y2]); A.append([0, 0, 0, x2, y2, 1, -y1 * x2, -y1 * y2])
        h = np.linalg.lstsq(np.array(A), a.reshape(-1), rcond=None)[0]
        Hm = np.append(h, 1).reshape(3, 3)
        q = np.c_[b, np.ones(len(b))] @ Hm.T
        res = np.linalg.norm(q[:, :2] / q[:, 2:] - a, axis=1).max()
        worst = max(worst, res)
    assert worst <= 3.5


def test_oracle_align_known_answer(oracle):
    """KAT: matchPairs.txt -> tran0.txt (BundleAdjustmentSparse, M/MosaicWithoutPos.cpp:6971-7202)."""
    pairs = _load_match_txt(); ref = _load_tran0()
    n_img = 20
    fixed = np.zeros(n_img, np.int32); fixed[0] = 1
    rc, out = oracle.align_affine(pairs, fixed)
    assert rc == 0
    assert np.array_equal(out[0], np.eye(3, dtype=np.float32).reshape(-1))
    got = out[1:, :8]
    assert np.allclose(got[:, :6], ref[:, :6], rtol=1e-4, atol=1e-2)
    assert np.abs(got[:, :6] - ref[:, :6]).max() < 1e-2            # fixture prints 6 significant digits
    # and against an independent double-precision least-squares solve
    rows, rhs = [], []
    for p in pairs:
        ia, xa, ya, fa, ib, xb, yb, fb = p
        ia, ib = int(ia), int(ib)
        rx = np.zeros(6 * 19); ry = np.zeros(6 * 19)
        bx = by = 0.0
        if fa == 0 and fb == 0:
            c = 6 * (ia - 1); rx[c + 0] = xa; rx[c + 1] = ya; rx[c + 4] = 1; ry[c + 2] = xa; ry[c + 3] = ya; ry[c + 5] = 1
            c = 6 * (ib - 1); rx[c + 0] = -xb; rx[c + 1] = -yb; rx[c + 4] = -1; ry[c + 2] = -xb; ry[c + 3] = -yb; ry[c + 5] = -1
        elif fa == 1 and fb == 0:                                  # image A is the fixed reference (identity)
            c = 6 * (ib - 1); rx[c + 0] = xb; rx[c + 1] = yb; rx[c + 4] = 1; ry[c + 2] = xb; ry[c + 3] = yb; ry[c + 5] = 1
            bx, by = xa, ya
        elif fa == 0 and fb == 1:
            c = 6 * (ia - 1); rx[c + 0] = xa; rx[c + 1] = ya; rx[c + 4] = 1; ry[c + 2] = xa; ry[c + 3] = ya; ry[c + 5] = 1
            bx, by = xb, yb
        else:
            continue
        rows.append(rx); rows.append(ry); rhs.append(bx); rhs.append(by)
    x = np.linalg.lstsq(np.array(rows), np.array(rhs), rcond=None)[0].reshape(19, 6)
    lsq = np.stack([x[:, 0], x[:, 1], x[:, 4], x[:, 2], x[:, 3], x[:, 5]], 1)
    assert np.allclose(got[:, :6], lsq, rtol=1e-6, atol=1e-4)


# ------------------------------------------------------------------------------------------------
# RANSAC: oracle restatement == the reference's own compiled code, bit for bit
# ------------------------------------------------------------------------------------------------
def test_oracle_ransac_vs_golden_reference_vectors(oracle):
    g = np.load(os.path.join(GOLD, "ransac_golden.npz"))
    for k in range(int(g["n_cases"][0])):
        seed, ok, ninl, rc = [int(v) for v in g[f"meta_{k}"]]
        o_ok, mask, H, n, st = oracle.ransac2d(g[f"xy1_{k}"], g[f"xy2_{k}"], 2.5, 1000, seed)
        assert (o_ok, n, int(st.rand_calls)) == (ok, ninl, rc), k
        assert np.array_equal(mask, g[f"mask_{k}"]), k
        assert np.array_equal(H.view(np.uint32), g[f"H_{k}"].view(np.uint32)) or np.array_equal(H, g[f"H_{k}"]), k


def test_oracle_ransac_vs_compiled_reference_live(oracle):
    if oracle.ref() is None:
        pytest.skip("oracle/_ref not built (no /root/reference on this box)")
    from imagemosaicing_b200 import synth
    rng = np.random.default_rng(99)
    for trial in range(25):
        w, h = [(4000, 3000), (1000, 750), (512, 512)][trial % 3]
        n = int(rng.integers(4, 397))
        xy1, xy2, _ = synth.make_candidates(rng, n, w, h, float(rng.uniform(0.2, 1.0)), 0.5)
        seed = int(rng.integers(0, 2 ** 32))
        r = oracle.ref_ransac2d(xy1, xy2, 2.5, 1000, seed)
        o = oracle.ransac2d(xy1, xy2, 2.5, 1000, seed)
        assert r[0] == o[0] and r[3] == o[3] and r[4] == o[4].rand_calls
        assert np.array_equal(r[1], o[1])
        assert np.array_equal(r[2].view(np.uint32), o[2].view(np.uint32)) or np.array_equal(r[2], o[2])
    # sub-functions: InverseMatrix, SolveHomographyMatrix, NonlinearLeastSquareProjection2
    for trial in range(20):
        M = rng.normal(0, 100, (8, 8)).astype(np.float32); M = (M @ M.T).astype(np.float32)
        for eps in (1e-6, 1e-20):
            rr, rd = oracle.inverse_matrix(M, eps, "ref"); orc, od = oracle.inverse_matrix(M, eps, "oracle")
            assert rr == orc and np.array_equal(rd.view(np.uint32), od.view(np.uint32))
        xy1, xy2, _ = synth.make_candidates(rng, 12, 1000, 750, 1.0, 0.3)
        hr = oracle.solve_homography(xy1, xy2, "ref"); ho = oracle.solve_homography(xy1, xy2, "oracle")
        assert np.array_equal(hr.view(np.uint32), ho.view(np.uint32))
        nr = oracle.nls_projection2(xy1, xy2, hr, 1e-10, "ref"); no = oracle.nls_projection2(xy1, xy2, hr, 1e-10, "oracle")
        assert np.array_equal(nr.view(np.uint32), no.view(np.uint32))


# ------------------------------------------------------------------------------------------------
# kernel math compiled for the host == oracle (per 4-tuple), sample stream
# ------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def hh():
    import subprocess
    d = os.path.join(ROOT, "tests", "host_harness")
    subprocess.check_call(["make", "-C", d, "-s"])
    return C.CDLL(os.path.join(d, "libhh.so"))


def test_device_ransac_math_on_host_matches_oracle(oracle, hh):
    from imagemosaicing_b200 import synth
    f32p = C.POINTER(C.c_float)
    rng = np.random.default_rng(3)
    n_slow = 0
    for trial in range(12):
        w, h = [(4000, 3000), (1000, 750), (512, 512)][trial % 3]
        xy1, xy2, _ = synth.make_candidates(rng, 396, w, h, float(rng.uniform(0.2, 1.0)), 0.5)
        if trial == 11:                                        # degenerate: collinear points -> slow path / gates
            xy2 = np.stack([np.linspace(0, 999, 396), np.linspace(0, 500, 396)], 1).astype(np.float32); xy1 = xy2 + 3
        for k in range(300):
            idx = rng.choice(396, 4, replace=False).astype(np.int32)
            kind, h_o, sup = oracle.ransac_eval_tuple(xy1, xy2, idx)
            a = [np.ascontiguousarray(v) for v in (xy1[idx, 0], xy1[idx, 1], xy2[idx, 0], xy2[idx, 1])]
            h_d = np.zeros(9, np.float32); slow = C.c_int(0)
            st = hh.hh_hypothesis(*[v.ctypes.data_as(f32p) for v in a], h_d.ctypes.data_as(f32p), C.byref(slow))
            n_slow += slow.value
            assert st == kind
            if kind != 0:
                assert np.array_equal(h_o.view(np.uint32), h_d.view(np.uint32)) or np.array_equal(h_o, h_d)
    assert n_slow > 0          # the generic (slow) path was exercised too


def test_sample_stream_jump_ahead(oracle, hh):
    """draw group g of the kernel == rand() calls 4g..4g+3 of the sequential MSVC LCG."""
    lib = oracle.lib()
    lib.orc_lcg_next.restype = C.c_uint32
    for seed, n in [(1, 396), (0xdeadbeef, 5), (123456789, 200)]:
        s = C.c_uint32(seed)
        seq = [lib.orc_lcg_next(C.byref(s)) % n for _ in range(4 * 600)]
        for g in (0, 1, 2, 17, 255, 599):
            idx = (C.c_int * 4)()
            valid = hh.hh_draw_group(C.c_uint32(seed), C.c_uint32(g), n, idx)
            want = seq[4 * g:4 * g + 4]
            assert list(idx) == want
            assert valid == (1 if len(set(want)) == 4 else 0)


# ------------------------------------------------------------------------------------------------
# oracle match / select against independent implementations
# ------------------------------------------------------------------------------------------------
def test_oracle_match_vs_numpy_and_cv2(oracle):
    from imagemosaicing_b200 import synth
    rng = np.random.default_rng(0)
    A = synth.sift_like_descriptors(rng, 300); B = synth.sift_like_descriptors(rng, 517)
    B[100] = B[3]; A[7] = B[3]                                   # tie: lowest index (3) must win
    idx, d2 = oracle.match_l2(A, B)
    D = ((A[:, None, :].astype(np.int64) - B[None, :, :].astype(np.int64)) ** 2).sum(2)
    assert np.array_equal(idx, D.argmin(1)) and np.array_equal(d2, D.min(1))
    assert idx[7] == 3
    try:
        import cv2
    except Exception:
        return
    m = cv2.BFMatcher(cv2.NORM_L2).match(A.astype(np.float32), B.astype(np.float32))
    cv_idx = np.array([x.trainIdx for x in m]); cv_d = np.array([x.distance for x in m], np.float32)
    assert np.array_equal(cv_idx, idx)
    assert np.array_equal(cv_d.view(np.uint32), np.sqrt(d2.astype(np.float32)).view(np.uint32))


def test_fast_cpu_matcher_equals_oracle(oracle):
    """bench.py's CPU arm uses a tuned (AVX2) matcher; it must return exactly what the plain restatement returns."""
    from imagemosaicing_b200 import synth
    rng = np.random.default_rng(5)
    for na, nb in [(1, 1), (3, 5), (257, 1031), (1024, 777)]:
        A = synth.sift_like_descriptors(rng, na); B = synth.sift_like_descriptors(rng, nb)
        if nb > 4:
            B[4] = B[1]; A[0] = B[1]
        a = oracle.match_l2(A, B); b = oracle.match_l2_fast(A, B)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    A = np.full((4, 128), 255, np.uint8); B = np.zeros((6, 128), np.uint8)      # largest distance 128 * 255^2
    a = oracle.match_l2(A, B); b = oracle.match_l2_fast(A, B)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[1][0] == 128 * 255 * 255


def _select_py(train, d2, kp1, kp2, w, h, gx=3, gy=3, max_num=400, frac=0.3):
    """Literal Python restatement of std::sort + SelectMatchPairs with the (d2, queryIdx) order."""
    order = sorted(range(len(train)), key=lambda q: (int(d2[q]), q))
    n_match = int(min(float(max_num), frac * len(train)))
    quota = int(np.float32(n_match) / np.float32(gx * gy))
    sx, sy = w // gx, h // gy
    label = {}
    out = []
    for q in order:
        nx = int(np.float32(kp1[q, 0]) / np.float32(sx)); ny = int(np.float32(kp1[q, 1]) / np.float32(sy))
        c = gx * ny + nx
        if label.get(c, 0) >= quota:
            continue
        out.append((q, int(train[q]))); label[c] = label.get(c, 0) + 1
    return out


def test_oracle_select_vs_literal_restatement(oracle):
    from imagemosaicing_b200 import synth
    rng = np.random.default_rng(5)
    for n, w, h in [(2048, 512, 512), (8192, 4000, 3000), (50, 1000, 750), (1334, 1000, 750)]:
        kp1 = synth.random_keypoints(rng, n, w, h); kp2 = synth.random_keypoints(rng, n, w, h)
        kp1[:3, 0] = w - 0.25
        train = rng.integers(0, n, n).astype(np.int32); d2 = rng.integers(0, 5000, n).astype(np.int32)
        x1, i1, x2, i2 = oracle.select(train, d2, kp1, kp2, w, h)
        want = _select_py(train, d2, kp1, kp2, w, h)
        assert [(int(a), int(b)) for a, b in zip(i1, i2)] == want
        assert np.array_equal(x1, kp1[i1]) and np.array_equal(x2, kp2[i2])
        assert len(i1) <= 400 + 16


# ------------------------------------------------------------------------------------------------
# product library: C-ABI surface and host-side stages (no GPU compute)
# ------------------------------------------------------------------------------------------------
def test_cabi_exports_every_declared_symbol():
    from imagemosaicing_b200 import _lib
    L = _lib.lib()
    hdr = open(os.path.join(ROOT, "include", "uavm.h")).read()
    names = sorted(set(re.findall(r"\b(uavm_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 35
    for nme in names:
        assert hasattr(L, nme), f"{nme} declared in include/uavm.h but not exported by libuavmosaic.so"
    assert C.sizeof(_lib.MatchPointPairs) == 40 and C.sizeof(_lib.SfPoint) == 12 and C.sizeof(_lib.DMatch) == 16
    assert C.sizeof(_lib.ImageTransform) == 40 and C.sizeof(_lib.ProjectMat) == 36


def test_no_gpu_means_loud_failure_not_fallback():
    """Without a CUDA device the context cannot be created (-2); there is no CPU path behind the ABI."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from imagemosaicing_b200 import api
    with pytest.raises(api.UavmError):
        api.Context(0)


def test_product_align_matches_oracle_and_golden(oracle):
    from imagemosaicing_b200 import _lib
    L = _lib.lib()
    pairs = _load_match_txt(); ref = _load_tran0()
    n = len(pairs)
    arr = (_lib.MatchPointPairs * n)()
    for k, p in enumerate(pairs):
        arr[k].ptA.x, arr[k].ptA.y, arr[k].ptA_i, arr[k].ptA_Fixed = float(p[1]), float(p[2]), int(p[0]), int(p[3])
        arr[k].ptB.x, arr[k].ptB.y, arr[k].ptB_i, arr[k].ptB_Fixed = float(p[5]), float(p[6]), int(p[4]), int(p[7])
    init = (_lib.ImageTransform * 20)(); out = (_lib.ImageTransform * 20)()
    for i in range(20):
        for t in range(9):
            init[i].h.m[t] = 1.0 if t in (0, 4, 8) else 0.0
        init[i].fixed = 1 if i == 0 else 0
    assert L.uavm_align_affine(arr, n, init, 20, 1, out) == 0
    got = np.array([[out[i].h.m[t] for t in range(9)] for i in range(20)], np.float32)
    assert np.abs(got[1:, :6] - ref[:, :6]).max() < 1e-2
    fixed = np.zeros(20, np.int32); fixed[0] = 1
    rc, o = oracle.align_affine(pairs, fixed)
    assert np.allclose(got, o, rtol=1e-6, atol=1e-5)
    # connectivity: all 20 images of the fixture are connected
    label = (C.c_int32 * 20)()
    assert L.uavm_connected_images(arr, n, 20, label) == 0
    assert list(label) == [1] * 20
    # two components: the larger one wins, the rest is labelled 0
    arr2 = (_lib.MatchPointPairs * 4)()
    for k, (a, b) in enumerate([(0, 1), (1, 2), (3, 4), (2, 0)]):
        arr2[k].ptA_i, arr2[k].ptB_i = a, b
    label6 = (C.c_int32 * 6)()
    assert L.uavm_connected_images(arr2, 4, 6, label6) == 0
    assert list(label6) == [1, 1, 1, 0, 0, 0]


def test_product_canvas_layout_matches_oracle(oracle):
    from imagemosaicing_b200 import api, synth
    rng = np.random.default_rng(8)
    for (w, h, n) in [(1000, 750, 5), (4000, 3000, 4), (512, 512, 2)]:
        T = [np.eye(3)]
        for k in range(1, n):
            Hk = synth.pair_homography(rng, w, h); Hk[2, :2] = 0
            T.append(T[-1] @ Hk)
        Hs = np.stack(T).astype(np.float32).reshape(n, 9)
        keep = np.ones(n, np.int32)
        if n > 3:
            keep[2] = 0
        c, ch = api.canvas_layout(Hs, keep, w, h)
        oc, och = oracle.canvas_layout(Hs, keep, w, h)
        assert (c.canvas_w, c.canvas_h, c.dgx, c.dgy) == (oc.canvas_w, oc.canvas_h, oc.dgx, oc.dgy)
        for k in range(n):
            a, b = ch[k], och[k]
            assert (a.keep, a.beg_x, a.beg_y, a.chip_w, a.chip_h, a.sx, a.sy) == (b.keep, b.beg_x, b.beg_y, b.chip_w, b.chip_h, b.sx, b.sy)
            assert list(a.quad) == list(b.quad) and list(a.inv) == list(b.inv)
    # implied canvas of the reference's shipped run (1647 x 2411, dG = (636.98, 1142.2))
    ref = _load_tran0()
    Hs = np.zeros((20, 9), np.float32); Hs[0] = np.eye(3).reshape(-1)
    Hs[1:, :8] = ref[:, :8]; Hs[1:, 8] = 1
    c, ch = api.canvas_layout(Hs, None, 1000, 750)
    assert (c.canvas_w, c.canvas_h) == (1647, 2411)
    assert abs(c.dgx - 636.98) < 0.05 and abs(c.dgy - 1142.2) < 0.05


def test_product_resample_by_overlap():
    """ResampleByOverlap (M/MosaicImage.cpp:2070-2201): golden keep flags produced by the reference's own
    compiled function (tests/golden/make_overlap_golden.py), plus live comparison when oracle/_ref exists."""
    from imagemosaicing_b200 import _lib
    from oracle import oracle as O
    L = _lib.lib()
    f32p = C.POINTER(C.c_float); i32p = C.POINTER(C.c_int32)

    def run(Hs, w, h, fn):
        Hs = np.ascontiguousarray(Hs, np.float32); keep = np.zeros(len(Hs), np.int32)
        assert fn(Hs.ctypes.data_as(f32p), len(Hs), w, h, C.c_float(0.7), keep.ctypes.data_as(i32p)) == 0
        return keep

    g = np.load(os.path.join(GOLD, "overlap_golden.npz"))
    dropped = 0
    for k in range(int(g["n_cases"][0])):
        w, h = [int(v) for v in g[f"wh_{k}"]]
        keep = run(g[f"H_{k}"], w, h, L.uavm_resample_by_overlap)
        assert np.array_equal(keep, g[f"keep_{k}"]), k
        dropped += int((keep == 0).sum())
        if O.ref() is not None:
            assert np.array_equal(run(g[f"H_{k}"], w, h, O.ref().ref_resample_by_overlap), keep)
    assert dropped > 20
    # 30 % steps: nothing dropped; the m[8] == 0 sentinel is ignored; the last image is always kept (:2198)
    T = lambda tx, ty: [1, 0, tx, 0, 1, ty, 0, 0, 1]
    assert list(run([T(0, 0), T(0, 225), T(0, 450), T(0, 675)], 1000, 750, L.uavm_resample_by_overlap)) == [1, 1, 1, 1]


# ------------------------------------------------------------------------------------------------
# K5 / K6: oracle restatement == the reference's own warp loop and FindMasksByDistMap
# ------------------------------------------------------------------------------------------------
def test_oracle_warp_and_masks_vs_golden_reference_vectors(oracle):
    """tests/golden/warp_golden.npz was produced by the reference's own code (M/MosaicImage.cpp:2350-2448,
    :1761-1881) compiled in place; the oracle's restatements must reproduce every byte."""
    g = np.load(os.path.join(GOLD, "warp_golden.npz"))
    for ci in range(int(g["n_cases"][0])):
        w, h, n = [int(v) for v in g[f"whn_{ci}"]]
        canvas, chips = oracle.canvas_layout(g[f"H_{ci}"], None, w, h)
        masks = []
        for k in range(n):
            px, m = oracle.warp_chip(g[f"img_{ci}_{k}"], canvas, chips[k])
            assert np.array_equal(px, g[f"chip_{ci}_{k}"]) and np.array_equal(m, g[f"mask_{ci}_{k}"]), (ci, k)
            masks.append(m)
            if oracle.ref() is not None:
                rpx, rm = oracle.ref_warp_chip(g[f"img_{ci}_{k}"], canvas, chips[k])
                assert np.array_equal(px, rpx) and np.array_equal(m, rm)
        seam = oracle.seam_masks(masks, [chips[k] for k in range(n)], canvas.canvas_w, canvas.canvas_h)
        for k in range(n):
            assert np.array_equal(seam[k], g[f"seam_{ci}_{k}"]), (ci, k)


def test_product_align_block_scale_equals_dense_oracle(oracle):
    """configs[2] scale: 200 images (10 strips x 20), 1194 unknowns.  The product solves the normal equations with a Cholesky
    restricted to the matrix envelope, on the decoupled x / y systems, with the matches of an image pair summed in local
    accumulators; the oracle factorises the full dense 6 (N - 1) matrix match by match.  Skipped terms are exact zeros; only the
    summation order inside a pair's run differs, so the float32 transforms agree to rounding (and the solve must be fast: the
    dense form takes ~0.4 s)."""
    import time
    from imagemosaicing_b200 import synth, _lib
    rows, cols, w, h = 10, 20, 4000, 3000
    poses = synth.block_poses(np.random.default_rng(1), rows, cols, w, h); n = rows * cols
    rng = np.random.default_rng(2)
    corners = np.array([[0, 0], [w - 1, 0], [w - 1, h - 1], [0, h - 1]], np.float64)
    boxes = []
    for T in poses:
        q = synth.apply_h(T, corners); boxes.append((q[:, 0].min(), q[:, 1].min(), q[:, 0].max(), q[:, 1].max()))
    rows_m = []
    for i in range(n):
        for j in range(i + 1, n):
            a, b = boxes[i], boxes[j]
            x0, y0, x1, y1 = max(a[0], b[0]), max(a[1], b[1]), min(a[2], b[2]), min(a[3], b[3])
            if x0 >= x1 or y0 >= y1:
                continue
            wp = np.stack([rng.uniform(x0, x1, 120), rng.uniform(y0, y1, 120)], 1)
            pi = synth.apply_h(np.linalg.inv(poses[i]), wp); pj = synth.apply_h(np.linalg.inv(poses[j]), wp)
            ok = (pi[:, 0] >= 0) & (pi[:, 0] <= w - 1) & (pi[:, 1] >= 0) & (pi[:, 1] <= h - 1) & (pj[:, 0] >= 0) & (pj[:, 0] <= w - 1) & (pj[:, 1] >= 0) & (pj[:, 1] <= h - 1)
            pi = pi[ok][:40] + rng.normal(0, 0.5, (min(40, int(ok.sum())), 2)); pj = pj[ok][:40]
            if len(pi) < 20:
                continue
            for k in range(len(pi)):
                rows_m.append([i, np.float32(pi[k, 0]), np.float32(pi[k, 1]), 1 if i == 0 else 0, j, np.float32(pj[k, 0]), np.float32(pj[k, 1]), 0])
    m = np.array(rows_m, np.float64); fixed = np.zeros(n, np.int32); fixed[0] = 1
    rc, T = oracle.align_affine(m, fixed)
    assert rc == 0
    arr = (_lib.MatchPointPairs * len(m))()
    for k, p in enumerate(m):
        arr[k].ptA.x, arr[k].ptA.y, arr[k].ptA_i, arr[k].ptA_Fixed = float(p[1]), float(p[2]), int(p[0]), int(p[3])
        arr[k].ptB.x, arr[k].ptB.y, arr[k].ptB_i, arr[k].ptB_Fixed = float(p[5]), float(p[6]), int(p[4]), int(p[7])
    init = (_lib.ImageTransform * n)(); out = (_lib.ImageTransform * n)()
    for i in range(n):
        for t in range(9):
            init[i].h.m[t] = 1.0 if t in (0, 4, 8) else 0.0
        init[i].fixed = 1 if i == 0 else 0
    t0 = time.perf_counter()
    assert _lib.lib().uavm_align_affine(arr, len(m), init, n, 1, out) == 0
    dt = time.perf_counter() - t0
    got = np.array([[out[i].h.m[t] for t in range(9)] for i in range(n)], np.float32)
    assert np.allclose(got, T, rtol=2e-6, atol=2e-5), np.abs(got - T).max()
    assert dt < 0.3, dt
    # the recovered block is close to the ground truth (the unconstrained affine objective drifts by a few px over 20 hops)
    err = [np.abs(synth.apply_h(np.linalg.inv(poses[0]) @ poses[k], corners) - synth.apply_h(got[k].reshape(3, 3).astype(np.float64), corners)).max() for k in range(n)]
    assert max(err) < 40.0


def test_header_is_c99_and_links_from_c(tmp_path):
    """include/uavm.h compiled as C99 by gcc (-pedantic) into a program that calls host-side entry points of the library."""
    import subprocess
    exe = str(tmp_path / "consumer")
    libdir = os.path.join(ROOT, "imagemosaicing_b200")
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "c_consumer", "consumer.c"), "-L", libdir, "-l:libuavmosaic.so", "-lm", "-Wl,-rpath," + libdir, "-o", exe]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0 and "c consumer ok" in r.stdout, r.stdout


def test_cpp_multi_gpu_host_compiles_and_links(tmp_path):
    """INTEGRATION.md section 2c as a real program (tests/c_consumer/multi_gpu_host.cpp: one thread per GPU, uavm_dist_*,
    uavm_pairbatch_allgather, uavm_global_align, uavm_canvas_set_rect / bind_root / gather): every call must match include/uavm.h
    (-Wall -Wextra -Werror) and link against the library; without a GPU it reports so and exits 0 (no CPU fallback)."""
    import subprocess
    exe = str(tmp_path / "multi_gpu_host")
    libdir = os.path.join(ROOT, "imagemosaicing_b200")
    cmd = ["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "c_consumer", "multi_gpu_host.cpp"), "-L", libdir, "-l:libuavmosaic.so", "-lpthread", "-Wl,-rpath," + libdir, "-o", exe]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([exe, "2"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=60)
        assert r.returncode == 0 and "no sm_100 device" in r.stdout, r.stdout


def test_constrained_alignment_variants_vs_oracle():
    """f3: uavm_align_affine_constrained / uavm_align_affine_rot (host C++) against the numpy restatement of
    BundleAdjustmentSparseConstraint / SparseAffineRotConstraint; and the point of the variants: on a chain of images the
    rotation-constrained solution is a near-similarity while the unconstrained affine fit drifts in scale / shear."""
    from imagemosaicing_b200 import synth, _lib
    from oracle import oracle_align as OA
    rng = np.random.default_rng(12)
    n, w, h = 8, 1000, 750
    poses = [np.eye(3)]
    for k in range(1, n):
        a = np.deg2rad(rng.uniform(-4, 4))
        poses.append(poses[-1] @ np.array([[np.cos(a), -np.sin(a), 420 + rng.uniform(-20, 20)], [np.sin(a), np.cos(a), rng.uniform(-30, 30)], [0, 0, 1]]))
    rows = []
    for i in range(n - 1):
        for j in range(i + 1, min(n, i + 3)):
            wp = synth.apply_h(poses[j], np.stack([rng.uniform(0, w - 1, 60), rng.uniform(0, h - 1, 60)], 1))
            pi = synth.apply_h(np.linalg.inv(poses[i]), wp); pj = synth.apply_h(np.linalg.inv(poses[j]), wp)
            ok = (pi[:, 0] >= 0) & (pi[:, 0] <= w - 1) & (pi[:, 1] >= 0) & (pi[:, 1] <= h - 1)
            pi = pi[ok] + rng.normal(0, 0.7, (int(ok.sum()), 2)); pj = pj[ok]
            for k in range(len(pi)):
                rows.append([i, np.float32(pi[k, 0]), np.float32(pi[k, 1]), 1 if i == 0 else 0, j, np.float32(pj[k, 0]), np.float32(pj[k, 1]), 0])
    m = np.array(rows, np.float64); fixed = np.zeros(n, np.int32); fixed[0] = 1
    T0 = np.tile(np.eye(3, dtype=np.float32).reshape(1, 9), (n, 1))
    arr = (_lib.MatchPointPairs * len(m))()
    for k, p in enumerate(m):
        arr[k].ptA.x, arr[k].ptA.y, arr[k].ptA_i, arr[k].ptA_Fixed = float(p[1]), float(p[2]), int(p[0]), int(p[3])
        arr[k].ptB.x, arr[k].ptB.y, arr[k].ptB_i, arr[k].ptB_Fixed = float(p[5]), float(p[6]), int(p[4]), int(p[7])
    init = (_lib.ImageTransform * n)(); out = (_lib.ImageTransform * n)()
    for i in range(n):
        for t in range(9):
            init[i].h.m[t] = 1.0 if t in (0, 4, 8) else 0.0
        init[i].fixed = int(fixed[i])
    lib = _lib.lib()
    get = lambda: np.array([[out[i].h.m[t] for t in range(9)] for i in range(n)], np.float32)
    assert lib.uavm_align_affine_constrained(arr, len(m), init, n, 1, out) == 0
    Tc = get()
    assert np.allclose(Tc, OA.align_affine_constrained(m, fixed, T0), rtol=1e-5, atol=1e-3)
    import ctypes as C
    assert lib.uavm_align_affine_rot(arr, len(m), init, n, 1, C.c_float(1.0), 10, out) == 0
    Tr = get()
    assert np.allclose(Tr, OA.align_affine_rot(m, fixed, T0, 1.0, 10), rtol=1e-4, atol=5e-3)
    assert lib.uavm_align_affine(arr, len(m), init, n, 1, out) == 0
    Tu = get()
    def nonsim(T):       # distance of the linear part from a rotation
        a, b, c, d = T[1:, 0], T[1:, 1], T[1:, 3], T[1:, 4]
        return np.max(np.abs(a * a + c * c - 1) + np.abs(b * b + d * d - 1) + np.abs(a * b + c * d))
    corners = np.array([[0, 0], [w - 1, 0], [w - 1, h - 1], [0, h - 1]], np.float64)
    err = lambda T: max(np.abs(synth.apply_h(poses[k], corners) - synth.apply_h(T[k].reshape(3, 3).astype(np.float64), corners)).max() for k in range(1, n))
    # with a strong constraint weight the solution is a near-rotation and does not drift more than the free affine fit
    assert lib.uavm_align_affine_rot(arr, len(m), init, n, 1, C.c_float(100.0), 10, out) == 0
    Ts = get()
    assert np.allclose(Ts, OA.align_affine_rot(m, fixed, T0, 100.0, 10), rtol=1e-4, atol=5e-3)
    print("non-similarity: free", nonsim(Tu), "rot w=1", nonsim(Tr), "rot w=100", nonsim(Ts), "corner err", err(Tu), err(Tr), err(Ts))
    assert nonsim(Ts) < 0.25 * nonsim(Tu)
    assert err(Ts) < err(Tu) + 1.0
