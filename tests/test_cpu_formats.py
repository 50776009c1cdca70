"""The reference's on-disk artefacts as first-class I/O (csrc/formats_host.cpp), against the reference's own
golden files: Release/feature_temp/matchPairs.match, matchPairs.txt and Release/tran0.txt (copied to tests/golden/ref_*),
and an OpenCV-written descriptor XML.  Host code only: runs without a GPU."""
import ctypes as C
import os
import numpy as np

from imagemosaicing_b200 import api, _lib as L

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_match_file_roundtrip_is_byte_identical(tmp_path):
    src = os.path.join(GOLD, "ref_matchPairs.match")
    m = api.read_match_file(src)
    assert len(m) == 5918 and C.sizeof(L.MatchPointPairs) == 40
    assert os.path.getsize(src) == 4 + 40 * len(m)
    out = str(tmp_path / "out.match")
    api.write_match_file(out, m)
    assert open(out, "rb").read() == open(src, "rb").read()
    # the reference writes nothing for an empty list (M/MosaicWithoutPos.cpp:4739)
    empty = str(tmp_path / "empty.match")
    api.write_match_file(empty, (L.MatchPointPairs * 0)())
    assert not os.path.exists(empty)


def test_match_text_reproduces_the_reference_file(tmp_path):
    """matchPairs.txt of the reference run = its matchPairs.match after fixing the reference image 0 (:4525-4556, :4575):
    writing our text form of that list must give the reference's file byte for byte (same %g formatting as ofstream << float)."""
    m = api.read_match_file(os.path.join(GOLD, "ref_matchPairs.match"))
    for r in m:
        if r.ptA_i == 0:
            r.ptA_Fixed = 1
        if r.ptB_i == 0:
            r.ptB_Fixed = 1
    out = str(tmp_path / "matchPairs.txt")
    api.write_match_text(out, m)
    ours = open(out).read().split("\n"); ref = open(os.path.join(GOLD, "ref_matchPairs.txt")).read().replace("\r", "").split("\n")
    assert len(ours) == len(ref)
    assert ours == ref
    back = api.read_match_text(out)
    assert len(back) == len(m)
    for a, b in zip(back[:200], m[:200]):
        assert (a.ptA_i, a.ptB_i, a.ptA_Fixed, a.ptB_Fixed) == (b.ptA_i, b.ptB_i, b.ptA_Fixed, b.ptB_Fixed)
        assert abs(a.ptA.x - b.ptA.x) <= 1e-3 * max(1.0, abs(b.ptA.x)) and abs(a.ptB.y - b.ptB.y) <= 1e-3 * max(1.0, abs(b.ptB.y))


def test_transform_file_roundtrip(tmp_path):
    src = os.path.join(GOLD, "ref_tran0.txt")
    T, fixed = api.read_transform_file(src)
    assert T.shape == (20, 9) and fixed[0] == 1 and not fixed[1:].any()
    assert np.array_equal(T[0], np.eye(3, dtype=np.float32).ravel())
    assert abs(T[1, 0] - 0.993179) < 1e-6 and abs(T[1, 5] + 100.622) < 1e-3 and T[1, 8] == 1.0
    out = str(tmp_path / "tran0.txt")
    api.write_transform_file(out, T, fixed)
    ours = [l.split() for l in open(out).read().strip().split("\n")]
    ref = [l.split() for l in open(src).read().replace("\r", "").strip().split("\n")]
    assert ours == ref                                    # 19 rows, m0..m7 + fixed, same digits
    # ImportTransform's format: count, 9 floats per image, image 0 fixed
    imp = str(tmp_path / "import.txt")
    with open(imp, "w") as f:
        f.write("2\n1 0 0 0 1 0 0 0 1\n0.5 0 3 0 0.5 4 0 0 1\n")
    T2, f2 = api.read_transform_file(imp, imported=True)
    assert T2.shape == (2, 9) and list(f2) == [1, 0] and T2[1, 2] == 3.0 and T2[1, 8] == 1.0


def test_feature_files(tmp_path):
    assert C.sizeof(L.KeyPoint) == 28                     # cv::KeyPoint of OpenCV 2.4
    rng = np.random.default_rng(3)
    n = 37
    kp = (L.KeyPoint * n)()
    for i in range(n):
        kp[i].x, kp[i].y, kp[i].size, kp[i].angle, kp[i].response = [float(v) for v in rng.random(5) * 100]
        kp[i].octave = int(rng.integers(0, 1 << 20)); kp[i].class_id = -1
    d = np.floor(rng.gamma(1.0, 30.0, (n, 128))).clip(0, 255).astype(np.float32)
    kpath, xpath = str(tmp_path / "keypoint_0.key"), str(tmp_path / "discriptor_0.xml")
    api.write_feature_files(kpath, xpath, kp, d)
    assert os.path.getsize(kpath) == 4 + 28 * n
    kp2, d2 = api.read_feature_files(kpath, xpath)
    assert bytes(kp2) == bytes(kp) and np.array_equal(d2, d)
    # a FileStorage file written by OpenCV itself (tests/golden/make_descriptor_xml_golden.py)
    r = C.c_int(0); c = C.c_int(0)
    gx = os.path.join(GOLD, "descriptor_cv2.xml")
    assert L.lib().uavm_descriptor_xml_size(gx.encode(), C.byref(r), C.byref(c)) == 0 and (r.value, c.value) == (7, 128)
    g = np.zeros((7, 128), np.float32)
    assert L.lib().uavm_descriptor_xml_read(gx.encode(), g.ctypes.data_as(L.f32p), g.size, C.byref(r), C.byref(c)) == 0
    assert np.array_equal(g, np.load(os.path.join(GOLD, "descriptor_cv2.npy")))
    try:
        import cv2
    except Exception:
        return
    fs = cv2.FileStorage(xpath, cv2.FILE_STORAGE_READ)    # and OpenCV reads what we write
    assert np.array_equal(fs.getNode("descriptor").mat(), d)


def test_io_errors_are_reported(tmp_path):
    n = C.c_int(0)
    assert L.lib().uavm_match_file_count(str(tmp_path / "missing.match").encode(), C.byref(n)) == -2
    short = str(tmp_path / "short.match")
    open(short, "wb").write(b"\x05\x00\x00\x00" + b"\x00" * 40)       # claims 5 records, holds 1
    arr = (L.MatchPointPairs * 8)()
    assert L.lib().uavm_match_file_read(short.encode(), arr, 8, C.byref(n)) == -2
    assert L.lib().uavm_match_file_read(os.path.join(GOLD, "ref_matchPairs.match").encode(), arr, 8, C.byref(n)) == -1 and n.value == 5918
