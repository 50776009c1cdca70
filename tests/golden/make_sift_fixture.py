#!/usr/bin/env python
"""Extracts SIFT features (cv2 4.13, SIFT_create(2000, 3, 0.01, 20) = the reference's parameters,
M/MosaicWithoutPos.cpp:4852) from the reference's 20 sample images (Release/test_data/DSC00004..23.JPG, the inputs of
M/mosaicing.cpp:18-49) and stores keypoint positions + u8 descriptors in tests/golden/ref_images_sift.npz.
Run in the build container (needs /root/reference and cv2); the GPU box only reads the .npz."""
import os
import numpy as np
import cv2
here = os.path.dirname(os.path.abspath(__file__))
src = "/root/reference/code/MosaicingCode/Release/test_data"
kps, descs, counts = [], [], []
sift = cv2.SIFT_create(2000, 3, 0.01, 20)
shape = None
for i in range(4, 24):
    img = cv2.imread(os.path.join(src, f"DSC{i:05d}.JPG"))
    assert img is not None
    shape = img.shape
    k, d = sift.detectAndCompute(img, None)
    k = k[:2000]; d = d[:2000]
    assert np.array_equal(d, np.rint(d)) and d.min() >= 0 and d.max() <= 255      # integer valued 0..255
    kps.append(np.array([p.pt for p in k], np.float32)); descs.append(d.astype(np.uint8)); counts.append(len(k))
np.savez_compressed(os.path.join(here, "ref_images_sift.npz"), counts=np.array(counts, np.int32), kp=np.concatenate(kps),
                    desc=np.concatenate(descs), width=shape[1], height=shape[0])
print("images", len(counts), "keypoints", sum(counts), "size", shape)
