#!/usr/bin/env python
"""f4: nvJPEG decode of the bench strip's frames (4000x3000, quality 90) into a canvas' BGR source pool — every backend the
library offers, frame after frame (uavm_canvas_set_image_jpeg) and as one batch (uavm_canvas_set_images_jpeg), with the
difference to the host decoder (cv2 / libjpeg-turbo).  Prints one JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, cv2
import bench
from imagemosaicing_b200 import api

NF = int(os.environ.get("NF", "49"))
W, H = bench.W, bench.H
descs, kps, Hs, T, base = bench.make_workload(0)
del descs, kps
frames = [np.roll(base, (37 * k) % H, axis=0) for k in range(NF)]
encs = [cv2.imencode(".jpg", f, [cv2.IMWRITE_JPEG_QUALITY, 90])[1] for f in frames]
t0 = time.perf_counter(); host = cv2.imdecode(encs[0], cv2.IMREAD_COLOR); host_ms = (time.perf_counter() - t0) * 1e3
ctx = api.Context(0, torch.cuda.current_stream())
keep = np.ones(NF, np.int32)
cv = api.Canvas(ctx, np.ascontiguousarray(T[:NF]), W, H, keep)
out = {"frames": NF, "frame": [W, H], "jpeg_bytes": int(sum(len(e) for e in encs)), "host_decode_ms_per_frame_1thread": host_ms, "runs": []}

def frame0():
    p, step = cv.image_ptr(0) if hasattr(cv, "image_ptr") else (None, None)
    return p, step

out["host_threads"] = os.cpu_count()
for backend in (3, 1, 0, 2):
    try:
        jp = api.Jpeg(ctx, backend)
    except api.UavmError as e:
        out["runs"].append({"backend": backend, "error": str(e)[:200]})
        continue
    hw = jp.hw_engines
    for mode in ("threads16", "threads8", "threads4", "nvjpeg_batched", "single"):
        if backend == 3 and mode != "nvjpeg_batched":
            continue
        if backend in (0, 2) and mode in ("threads8", "threads4"):
            continue
        rec = {"backend": backend, "mode": mode, "hw_engines": hw}
        try:
            if mode.startswith("threads"):
                jp.set_threads(int(mode[7:]))
            elif mode == "nvjpeg_batched":
                jp.set_threads(-1)
            def step():
                if mode == "single":
                    for k in range(NF):
                        jp.set_canvas_image(cv, k, encs[k])
                else:
                    jp.set_canvas_images(cv, 0, encs)
            step(); ctx.sync()
            t0 = time.perf_counter()
            for _ in range(2):
                step()
            ctx.sync()
            rec["ms_per_frame"] = (time.perf_counter() - t0) / 2 / NF * 1e3
            worst = 0; mean = 0.0
            for k in (0, 1, NF // 2, NF - 1):
                hostk = cv2.imdecode(encs[k], cv2.IMREAD_COLOR)
                d = np.abs(cv.source_frame(k).cpu().numpy().astype(np.int16) - hostk.astype(np.int16))
                worst = max(worst, int(d.max())); mean = max(mean, float(d.mean()))
            rec["vs_host_decoder"] = {"max": worst, "mean": mean, "frames_checked": 4}
        except api.UavmError as e:
            rec["error"] = str(e)[:200]
        out["runs"].append(rec)
    jp.close()
print(json.dumps(out))
