"""GPU parity tests of K5 (chip warp) and K6 (seam masks) against the CPU oracle, through the C ABI.
Bar: chip bytes and masks identical (the kernels reproduce the reference's float expression order)."""
import numpy as np
import pytest

from imagemosaicing_b200 import api, synth

pytestmark = pytest.mark.gpu


def _transforms(rng, n, w, h, projective=False):
    """Chain of transforms image k -> mosaic frame (image 0 = identity), affine like BundleAdjustmentSparse's output."""
    T = [np.eye(3)]
    for k in range(1, n):
        H = synth.pair_homography(rng, w, h, overlap=(0.55, 0.8))
        if not projective:
            H[2, :2] = 0
        T.append(T[-1] @ H)
    out = np.stack([t / t[2, 2] for t in T]).astype(np.float32).reshape(n, 9)
    return out


def _compare_layout(canvas, chips, o_canvas, o_chips, n):
    assert (canvas.canvas_w, canvas.canvas_h) == (o_canvas.canvas_w, o_canvas.canvas_h)
    assert canvas.dgx == o_canvas.dgx and canvas.dgy == o_canvas.dgy
    for k in range(n):
        a, b = chips[k], o_chips[k]
        assert (a.keep, a.beg_x, a.beg_y, a.chip_w, a.chip_h) == (b.keep, b.beg_x, b.beg_y, b.chip_w, b.chip_h)
        assert a.sx == b.sx and a.sy == b.sy
        assert list(a.quad) == list(b.quad) and list(a.inv) == list(b.inv)


@pytest.mark.parametrize("w,h,n,projective", [(512, 512, 2, False), (1000, 750, 4, False), (640, 480, 3, True), (333, 257, 3, False)])
def test_warp_and_masks_parity(ctx, oracle, w, h, n, projective):
    rng = np.random.default_rng(w * 7 + n)
    H = _transforms(rng, n, w, h, projective)
    imgs = [synth.texture_image(rng, w, h, 8) for _ in range(n)]
    cv = api.Canvas(ctx, H, w, h)
    o_canvas, o_chips = oracle.canvas_layout(H, None, w, h)
    _compare_layout(cv.layout, cv.chips, o_canvas, o_chips, n)
    for k in range(n):
        cv.set_image(k, imgs[k])
    cv.warp()
    o_masks = []
    for k in range(n):
        px, mask = cv.chip(k)
        o_px, o_mask = oracle.warp_chip(imgs[k], o_canvas, o_chips[k])
        assert np.array_equal(mask, o_mask)
        assert np.array_equal(px, o_px), f"chip {k}: {(px != o_px).sum()} differing bytes"
        o_masks.append(o_mask)
    cv.seam_masks()
    o_seam = oracle.seam_masks(o_masks, [o_chips[k] for k in range(n)], o_canvas.canvas_w, o_canvas.canvas_h)
    for k in range(n):
        _, mask = cv.chip(k)
        assert np.array_equal(mask, o_seam[k]), f"seam mask {k}: {(mask != o_seam[k]).sum()} differing pixels"
    # every covered canvas pixel has exactly one owner at most
    own = np.zeros((o_canvas.canvas_h, o_canvas.canvas_w), np.int32)
    for k in range(n):
        c = o_chips[k]
        own[c.beg_y:c.beg_y + c.chip_h, c.beg_x:c.beg_x + c.chip_w] += (o_seam[k] > 0)
    assert own.max() <= 1


def test_warp_skip_flags_and_full_frame(ctx, oracle):
    """keep flags / m[8]==0 sentinel (M/MosaicImage.cpp:2243,2311) and one full-size 4000x3000 frame."""
    rng = np.random.default_rng(3)
    w, h = 4000, 3000
    H = _transforms(rng, 3, w, h)
    H[1, 8] = 0.0                      # "skip me"
    keep = np.array([1, 1, 1], np.int32)
    img = synth.texture_image(rng, w, h, 4)
    cv = api.Canvas(ctx, H, w, h, keep)
    o_canvas, o_chips = oracle.canvas_layout(H, keep, w, h)
    _compare_layout(cv.layout, cv.chips, o_canvas, o_chips, 3)
    assert cv.chips[1].keep == 0
    for k in (0, 2):
        cv.set_image(k, img)
    cv.warp()
    for k in (0, 2):
        px, mask = cv.chip(k)
        o_px, o_mask = oracle.warp_chip(img, o_canvas, o_chips[k])
        assert np.array_equal(mask, o_mask) and np.array_equal(px, o_px)


def _affine(angle_deg, scale, tx, ty):
    a = np.deg2rad(angle_deg)
    return np.array([[scale * np.cos(a), -scale * np.sin(a), tx], [scale * np.sin(a), scale * np.cos(a), ty], [0, 0, 1.0]])


@pytest.mark.parametrize("w,h,angle,scale", [(640, 480, 33.0, 1.0), (640, 480, -80.0, 1.3), (512, 384, 3.0, 0.45), (640, 480, 2.0, 2.6),
                                              (514, 390, 5.0, 1.0), (40, 30, 10.0, 1.0), (700, 500, 180.0, 1.0)])
def test_warp_affine_staging_edges(ctx, oracle, w, h, angle, scale):
    """The packed affine kernel stages the source footprint of every 128 x 32 chip tile in shared memory (TMA boxes of
    160 x 64 px at most).  Strong rotations / reductions make footprints that do not fit (direct-load path), enlargements
    make tiny ones, an image width that is not a multiple of 4 disables the tensor map, and the rotated chips have tiles
    entirely outside the source (zero fill).  All paths must stay byte-identical to the reference loop."""
    rng = np.random.default_rng(int(abs(angle) * 10) + w)
    H = np.stack([np.eye(3), _affine(angle, scale, 0.31 * w, -0.22 * h)]).astype(np.float32).reshape(2, 9)
    imgs = [synth.texture_image(rng, w, h, 8) for _ in range(2)]
    cv = api.Canvas(ctx, H, w, h)
    o_canvas, o_chips = oracle.canvas_layout(H, None, w, h)
    _compare_layout(cv.layout, cv.chips, o_canvas, o_chips, 2)
    for k in range(2):
        cv.set_image(k, imgs[k])
    cv.warp()
    for k in range(2):
        px, mask = cv.chip(k)
        o_px, o_mask = oracle.warp_chip(imgs[k], o_canvas, o_chips[k])
        assert np.array_equal(mask, o_mask), f"chip {k}: {(mask != o_mask).sum()} differing mask bytes"
        assert np.array_equal(px, o_px), f"chip {k}: {(px != o_px).sum()} differing bytes"


def test_warp_range_equals_warp(ctx):
    """uavm_canvas_warp_range over consecutive groups == one uavm_canvas_warp (streaming callers warp group by group)."""
    rng = np.random.default_rng(11)
    w, h, n = 640, 480, 5
    H = _transforms(rng, n, w, h)
    imgs = [synth.texture_image(rng, w, h, 6) for _ in range(n)]
    a = api.Canvas(ctx, H, w, h); b = api.Canvas(ctx, H, w, h)
    for k in range(n):
        a.set_image(k, imgs[k])
    a.warp()
    for g0 in range(0, n, 2):
        for k in range(g0, min(g0 + 2, n)):
            b.set_image(k, imgs[k])
        b.warp(g0, min(2, n - g0))
    for k in range(n):
        pa, ma = a.chip(k); pb_, mb = b.chip(k)
        assert np.array_equal(pa, pb_) and np.array_equal(ma, mb)


@pytest.mark.parametrize("w,h,n,bands", [(320, 240, 3, 5), (500, 375, 4, 5), (257, 190, 2, 3), (640, 480, 5, 5)])
def test_multiband_blend_parity(ctx, oracle, w, h, n, bands):
    """K7 against the oracle's restatement of MultiBandBlender and against the cv2 4.13 blender itself (the proxy for the
    reference's OpenCV 2.4.0 binary): both bit-exact — the f32 weight pyrDown follows OpenCV's operation order."""
    rng = np.random.default_rng(w + n)
    H = _transforms(rng, n, w, h, False)
    imgs = [synth.texture_image(rng, w, h, 5) for _ in range(n)]
    cv = api.Canvas(ctx, H, w, h)
    for k in range(n):
        cv.set_image(k, imgs[k])
    cv.warp(); cv.seam_masks(); cv.blend(bands)
    out, om = cv.result()
    o_canvas, o_chips = oracle.canvas_layout(H, None, w, h)
    chips, masks = [], []
    for k in range(n):
        px, m = oracle.warp_chip(imgs[k], o_canvas, o_chips[k]); chips.append(px); masks.append(m)
    seam = oracle.seam_masks(masks, [o_chips[k] for k in range(n)], o_canvas.canvas_w, o_canvas.canvas_h)
    tls = [(o_chips[k].beg_x, o_chips[k].beg_y) for k in range(n)]
    o_out, o_om = oracle.multiband_blend(chips, seam, tls, o_canvas.canvas_w, o_canvas.canvas_h, bands)
    assert np.array_equal(om, o_om)
    assert np.array_equal(out, o_out), f"{(out != o_out).sum()} differing bytes"
    try:
        import cv2
    except Exception:
        return
    b = cv2.detail_MultiBandBlender(0, bands)
    b.prepare((0, 0, o_canvas.canvas_w, o_canvas.canvas_h))
    for k in range(n):
        b.feed(chips[k].astype(np.int16), seam[k], tls[k])
    rs, rm = b.blend(None, None)
    r8 = np.clip(rs, 0, 255).astype(np.uint8)
    assert np.array_equal(om, rm)
    assert np.array_equal(out, r8), f"{(out != r8).sum()} bytes differ from cv2.detail_MultiBandBlender"


def test_blend_without_seam_masks_parity(ctx, oracle):
    """The blender fed with the plain validity masks (no FindMasksByDistMap): every overlapping chip contributes to a pixel,
    so the f32 weight sums depend on the image order — the canvas kernels gather in index order like feed()."""
    rng = np.random.default_rng(77)
    w, h, n = 400, 300, 4
    H = _transforms(rng, n, w, h, False)
    imgs = [synth.texture_image(rng, w, h, 5) for _ in range(n)]
    cv = api.Canvas(ctx, H, w, h)
    for k in range(n):
        cv.set_image(k, imgs[k])
    cv.warp(); cv.blend(4)
    out, om = cv.result()
    o_canvas, o_chips = oracle.canvas_layout(H, None, w, h)
    chips, masks = [], []
    for k in range(n):
        px, m = oracle.warp_chip(imgs[k], o_canvas, o_chips[k]); chips.append(px); masks.append(m)
    tls = [(o_chips[k].beg_x, o_chips[k].beg_y) for k in range(n)]
    o_out, o_om = oracle.multiband_blend(chips, masks, tls, o_canvas.canvas_w, o_canvas.canvas_h, 4)
    assert np.array_equal(om, o_om)
    assert np.array_equal(out, o_out), f"{(out != o_out).sum()} differing bytes"


def test_full_size_seam_and_blend_parity(ctx, oracle):
    """BASELINE-size frames: three overlapping 4000 x 3000 chips through K5 + K6 + K7 against the oracle (seam masks byte
    for byte, blended canvas byte for byte)."""
    rng = np.random.default_rng(4000)
    w, h, n = 4000, 3000, 3
    H = _transforms(rng, n, w, h, False)
    base = synth.texture_image(rng, w, h, 4)
    imgs = [np.roll(base, 311 * k, axis=0) for k in range(n)]
    cv = api.Canvas(ctx, H, w, h)
    for k in range(n):
        cv.set_image(k, imgs[k])
    cv.warp()
    o_canvas, o_chips = oracle.canvas_layout(H, None, w, h)
    chips, masks = [], []
    for k in range(n):
        px, m = cv.chip(k)                                   # chips are pinned to the oracle by the K5 tests
        chips.append(px); masks.append(m)
    cv.seam_masks()
    seam = oracle.seam_masks(masks, [o_chips[k] for k in range(n)], o_canvas.canvas_w, o_canvas.canvas_h)
    for k in range(n):
        _, m = cv.chip(k)
        assert np.array_equal(m, seam[k]), f"seam mask {k}: {(m != seam[k]).sum()} differing pixels"
    cv.blend(5)
    out, om = cv.result()
    tls = [(o_chips[k].beg_x, o_chips[k].beg_y) for k in range(n)]
    o_out, o_om = oracle.multiband_blend(chips, seam, tls, o_canvas.canvas_w, o_canvas.canvas_h, 5)
    assert np.array_equal(om, o_om)
    assert np.array_equal(out, o_out), f"{(out != o_out).sum()} differing bytes"


def _grid_transforms(rng, cols, rows, w, h):
    """configs[4]-shaped block: cols x rows frames on a (0.375 w, 0.649 h) pitch with +-2 deg / +-2 % jitter."""
    T = np.zeros((cols * rows, 9), np.float32)
    for r in range(rows):
        for c in range(cols):
            k = r * cols + c
            a = 0.0 if k == 0 else np.deg2rad(rng.uniform(-2, 2)); s = 1.0 if k == 0 else rng.uniform(0.98, 1.02)
            T[k] = [s * np.cos(a), -s * np.sin(a), c * 0.375 * w, s * np.sin(a), s * np.cos(a), r * 0.649 * h, 0, 0, 1]
    return T


@pytest.mark.parametrize("world,grid2d", [(8, False), (8, True), (4, True), (3, False)])
def test_sharded_blend_equals_unsharded_2d_grid(ctx, world, grid2d):
    """Tile/shard equivalence on a configs[4]-shaped 2-D block (row AND column neighbours): the canvas computed as 8 row bands
    or as a 4 x 2 grid of rectangles — what the ranks of a multi-GPU run do — equals the unsharded blend byte for byte."""
    from imagemosaicing_b200 import dist as D
    rng = np.random.default_rng(5)
    w, h, cols, rows = 400, 300, 6, 5
    H = _grid_transforms(rng, cols, rows, w, h)
    n = cols * rows
    imgs = [synth.texture_image(rng, w, h, 5) for _ in range(4)]
    cv = api.Canvas(ctx, H, w, h)
    for k in range(n):
        cv.set_image(k, imgs[k % 4])
    cv.warp(); cv.seam_masks(); cv.blend(5)
    full, full_mask = cv.result()
    cw, ch = cv.layout.canvas_w, cv.layout.canvas_h
    rects = D.canvas_grid(cw, ch, world) if grid2d else [(0, y0, cw, y1) for (y0, y1) in D.canvas_bands(ch, world)]
    tiled = np.zeros_like(full); tiled_mask = np.zeros_like(full_mask)
    n_inactive = 0
    for (x0, y0, x1, y1) in rects:
        cvb = api.Canvas(ctx, H, w, h)
        cvb.set_rect(x0, y0, x1, y1)
        for k in range(n):
            if cvb.is_active(k):
                cvb.set_image(k, imgs[k % 4])
            else:
                n_inactive += 1
        cvb.warp(); cvb.seam_masks(); cvb.blend(5)
        out, om = cvb.result()
        tiled[y0:y1, x0:x1] = out[y0:y1, x0:x1]; tiled_mask[y0:y1, x0:x1] = om[y0:y1, x0:x1]
        piece = np.zeros((y1 - y0, x1 - x0, 3), np.uint8)
        cvb.copy_result_rect(x0, y0, x1, y1, piece)
        assert np.array_equal(piece, out[y0:y1, x0:x1])
        cvb.close()
    assert np.array_equal(tiled_mask, full_mask)
    assert np.array_equal(tiled, full), f"{(tiled != full).sum()} differing bytes"
    assert n_inactive > 0


def test_banded_blend_equals_untiled(ctx):
    """Tile/shard equivalence: computing the canvas in horizontal bands, as the ranks of a multi-GPU run do,
    reproduces the untiled blend byte for byte."""
    from imagemosaicing_b200 import dist as D
    rng = np.random.default_rng(21)
    w, h, n = 320, 240, 14
    descs, kps, Hs = synth.make_strip(n, w, h, 16, seed=3)
    T = [np.eye(3)]
    for Hk in Hs:
        A = Hk / Hk[2, 2]; A[2, :2] = 0
        T.append(T[-1] @ A)
    H = np.stack(T).astype(np.float32).reshape(n, 9)
    imgs = [synth.texture_image(rng, w, h, 5) for _ in range(n)]
    cv = api.Canvas(ctx, H, w, h)
    for k in range(n):
        cv.set_image(k, imgs[k])
    cv.warp(); cv.seam_masks(); cv.blend(5)
    full, full_mask = cv.result()
    ch = cv.layout.canvas_h
    assert ch > 3 * 160
    for world in (2, 3):
        tiled = np.zeros_like(full); tiled_mask = np.zeros_like(full_mask)
        n_inactive = 0
        for (y0, y1) in D.canvas_bands(ch, world):
            cvb = api.Canvas(ctx, H, w, h)
            cvb.set_band(y0, y1)
            for k in range(n):
                if cvb.is_active(k):
                    cvb.set_image(k, imgs[k])
                else:
                    n_inactive += 1
            cvb.warp(); cvb.seam_masks(); cvb.blend(5)
            out, om = cvb.result()
            tiled[y0:y1] = out[y0:y1]; tiled_mask[y0:y1] = om[y0:y1]
            cvb.close()
        assert np.array_equal(tiled_mask, full_mask)
        assert np.array_equal(tiled, full), f"world {world}: {(tiled != full).sum()} differing bytes"
        assert n_inactive > 0            # bands really skip chips that cannot touch them


def test_warp_for_blend_equals_full_warp(ctx):
    """K6 before K5: warping only what the blend reads gives the same mosaic byte for byte (unsharded and in a rectangle)."""
    rng = np.random.default_rng(31)
    w, h, cols, rows = 400, 300, 4, 3
    H = _grid_transforms(rng, cols, rows, w, h); n = cols * rows
    imgs = [synth.texture_image(rng, w, h, 5) for _ in range(3)]
    a = api.Canvas(ctx, H, w, h)
    for k in range(n):
        a.set_image(k, imgs[k % 3])
    a.warp(); a.seam_masks(); a.blend(5)
    full, fm = a.result()
    b = api.Canvas(ctx, H, w, h)
    for k in range(n):
        b.set_image(k, imgs[k % 3])
    b.seam_masks(); b.warp_for_blend(); b.blend(5)
    out, om = b.result()
    assert np.array_equal(om, fm) and np.array_equal(out, full)
    b.blend(3)                                            # fewer bands read less: still covered
    a.blend(3)
    assert np.array_equal(b.result()[0], a.result()[0])
    cw, ch = b.layout.canvas_w, b.layout.canvas_h
    x0, y0, x1, y1 = (cw // 3) & ~1, (ch // 4) & ~1, (2 * cw // 3) & ~1, ch
    c = api.Canvas(ctx, H, w, h)
    c.set_rect(x0, y0, x1, y1)
    for k in range(n):
        if c.is_active(k):
            c.set_image(k, imgs[k % 3])
    c.seam_masks(); c.warp_for_blend(); c.blend(5)
    part, _ = c.result()
    assert np.array_equal(part[y0:y1, x0:x1], full[y0:y1, x0:x1])
