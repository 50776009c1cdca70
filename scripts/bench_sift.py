#!/usr/bin/env python
"""GPU SIFT (csrc/sift.cu) on one BASELINE-size frame (4000x3000) against cv2 4.13 SIFT_create(2000, 3, 0.01, 20) on the host:
ms per frame, keypoints found, and the agreement of the two keypoint sets."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, cv2
from imagemosaicing_b200 import api, synth

w, h = 4000, 3000
rng = np.random.default_rng(5)
img = synth.texture_image(rng, w, h, 7)
n = cv2.GaussianBlur(rng.normal(0, 1, (h, w)).astype(np.float32), (0, 0), 1.3)
img = np.clip(img.astype(np.float32) * 0.6 + 50 + 260 * n[..., None], 0, 255).astype(np.uint8)
ctx = api.Context(0, torch.cuda.current_stream())
s = api.Sift(ctx, w, h)
d_img = torch.from_numpy(img).cuda()
kp, desc = s.detect_and_compute(d_img)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    kp, desc = s.detect_and_compute(d_img)
gpu_ms = (time.perf_counter() - t0) / 5 * 1e3
t0 = time.perf_counter()
for _ in range(3):
    kp_h, desc_h = s.detect_and_compute(img)
gpu_host_ms = (time.perf_counter() - t0) / 3 * 1e3
ref = cv2.SIFT_create(2000, 3, 0.01, 20)
cv2.setNumThreads(os.cpu_count())
t0 = time.perf_counter(); kc, dc = ref.detectAndCompute(img, None); cv_ms = (time.perf_counter() - t0) * 1e3
pc = np.array([k.pt for k in kc], np.float32)
pg = np.stack([kp["x"], kp["y"]], 1)
# nearest GPU keypoint for each cv2 keypoint (coarse grid hash would be faster; 2000 x 2000 is fine)
d = np.sqrt(((pc[:, None, :] - pg[None, :, :]) ** 2).sum(-1)).min(1)
print(json.dumps({"frame": [w, h], "gpu_ms_per_frame_device_input": gpu_ms, "gpu_ms_per_frame_host_input": gpu_host_ms, "cv2_ms_per_frame_all_host_threads": cv_ms,
                  "host_threads": os.cpu_count(), "keypoints_gpu": int(len(kp)), "keypoints_cv2": int(len(kc)),
                  "cv2_keypoints_with_gpu_keypoint_within_0.05px": float((d < 0.05).mean())}))
