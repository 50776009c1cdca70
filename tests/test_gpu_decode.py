"""f4: JPEG frames decoded on the device (csrc/decode.cu, nvJPEG) straight into the canvas' source pool."""
import numpy as np
import pytest

from imagemosaicing_b200 import api, synth

pytestmark = pytest.mark.gpu
cv2 = pytest.importorskip("cv2")


@pytest.mark.parametrize("backend", [0, 2])
def test_jpeg_decode_matches_host_decoder(ctx, backend):
    """nvJPEG against libjpeg-turbo (cv2.imdecode) on the same stream: different IDCT / chroma upsampling arithmetic, so the
    bar is a few grey levels, not bit equality (the reference's libjpeg differs from both in the same way)."""
    import torch
    rng = np.random.default_rng(1)
    w, h = 640, 480
    img = synth.texture_image(rng, w, h, 6)
    ok, enc = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, 92])
    assert ok
    host = cv2.imdecode(enc, cv2.IMREAD_COLOR)
    try:
        j = api.Jpeg(ctx, backend)
    except api.UavmError as e:
        pytest.skip(f"nvJPEG backend {backend} unavailable: {e}")
    out = torch.zeros((h, w, 3), dtype=torch.uint8, device="cuda")
    j.decode(enc, out)
    ctx.sync(); torch.cuda.synchronize()
    d = np.abs(out.cpu().numpy().astype(np.int32) - host.astype(np.int32))
    print(f"nvjpeg backend {backend}: |diff| max {d.max()} mean {d.mean():.3f} p99.9 {np.percentile(d, 99.9):.0f}")
    assert d.mean() < 1.5 and d.max() <= 12, (d.max(), d.mean())
    j.close()


def test_canvas_from_jpeg_frames(ctx, oracle):
    """Frames handed over as JPEG bytes: decoded into the BGR pool, warped; the chips equal the oracle's warp of the SAME decoded
    pixels byte for byte (the decoder's output is the input of the bit-exact path)."""
    import torch
    rng = np.random.default_rng(2)
    w, h, n = 640, 480, 3
    T = [np.eye(3)]
    for k in range(1, n):
        Hk = synth.pair_homography(rng, w, h, overlap=(0.55, 0.8)); Hk[2, :2] = 0
        T.append(T[-1] @ Hk)
    H = np.stack(T).astype(np.float32).reshape(n, 9)
    cv = api.Canvas(ctx, H, w, h)
    assert cv.source_layout == 3
    j = api.Jpeg(ctx, 0)
    decoded = []
    for k in range(n):
        ok, enc = cv2.imencode(".jpg", synth.texture_image(rng, w, h, 6), [cv2.IMWRITE_JPEG_QUALITY, 90])
        j.set_canvas_image(cv, k, enc)
        buf = torch.zeros((h, w, 3), dtype=torch.uint8, device="cuda")
        j.decode(enc, buf)
        ctx.sync()                                         # the decode runs on the library's stream
        decoded.append(buf.cpu().numpy())
    cv.warp()
    o_canvas, o_chips = oracle.canvas_layout(H, None, w, h)
    for k in range(n):
        px, mask = cv.chip(k)
        o_px, o_mask = oracle.warp_chip(decoded[k], o_canvas, o_chips[k])
        assert np.array_equal(mask, o_mask) and np.array_equal(px, o_px)
    j.close()


@pytest.mark.parametrize("threads", [3, -1])
def test_jpeg_batch_equals_frame_by_frame(ctx, threads):
    """uavm_canvas_set_images_jpeg (frames spread over host threads, one decoder lane and stream each; -1: nvJPEG's own batched
    decoder) leaves the same pixels in the source pool as decoding frame after frame, back to back without a sync in between
    (a decoder state is only reused after its previous decode has finished on the device)."""
    rng = np.random.default_rng(3)
    w, h, n = 640, 480, 7
    H = np.tile(np.eye(3, dtype=np.float32).reshape(1, 9), (n, 1))
    cv = api.Canvas(ctx, H, w, h)
    encs = [cv2.imencode(".jpg", synth.texture_image(rng, w, h, 6), [cv2.IMWRITE_JPEG_QUALITY, 90])[1] for _ in range(n)]
    j = api.Jpeg(ctx, 1)
    for k in range(n):
        j.set_canvas_image(cv, k, encs[k])
    ctx.sync()
    single = [cv.source_frame(k).cpu().numpy().copy() for k in range(n)]
    for k in range(n):
        host = cv2.imdecode(encs[k], cv2.IMREAD_COLOR)
        d = np.abs(single[k].astype(np.int32) - host.astype(np.int32))
        assert d.mean() < 1.5 and d.max() <= 12, (k, d.max(), d.mean())
    for k in range(n):
        cv.source_frame(k).zero_()
    import torch
    torch.cuda.synchronize()
    j.set_threads(threads)
    try:
        j.set_canvas_images(cv, 0, encs)
    except api.UavmError as e:
        if threads == -1:
            pytest.skip(f"nvJPEG batched decoder unavailable: {e}")
        raise
    ctx.sync()
    for k in range(n):
        got = cv.source_frame(k).cpu().numpy()
        if threads == -1:                                  # nvJPEG's batched decoder may pick another backend: decoder tolerance
            d = np.abs(got.astype(np.int32) - single[k].astype(np.int32))
            assert d.max() <= 12, (k, d.max())
        else:
            assert np.array_equal(got, single[k]), k
    j.close()
