#!/usr/bin/env python
"""Writes tests/golden/descriptor_cv2.xml with OpenCV's own FileStorage (cv2 4.13 here; the reference used 2.4.0's,
M/MosaicWithoutPos.cpp:4687-4689: fs << "descriptor" << descriptors) and the matrix it holds as descriptor_cv2.npy."""
import os
import numpy as np
import cv2
here = os.path.dirname(os.path.abspath(__file__))
rng = np.random.default_rng(24)
d = np.floor(rng.gamma(1.0, 30.0, (7, 128))).clip(0, 255).astype(np.float32)
fs = cv2.FileStorage(os.path.join(here, "descriptor_cv2.xml"), cv2.FILE_STORAGE_WRITE)
fs.write("descriptor", d)
fs.release()
np.save(os.path.join(here, "descriptor_cv2.npy"), d)
