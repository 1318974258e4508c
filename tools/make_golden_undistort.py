"""Golden vectors for Frame::UndistortKeyPoints / ComputeImageBounds: cv2.undistortPoints (OpenCV 4.13 wheel of the build container)
on seeded points.  Run once in the container (needs cv2); the fixture tests/golden/undistort.npz is committed."""
import os

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rng = np.random.default_rng(7)
K4 = np.array([558.4684, 560.0944, 326.7993, 262.9017], np.float32)          # Dual-LenaCV.yaml camera 1
K = np.array([[K4[0], 0, K4[2]], [0, K4[1], K4[3]], [0, 0, 1]], np.float32)
pts = np.concatenate([rng.uniform([-5, -5], [645, 485], (600, 2)), [[0, 0], [640, 0], [0, 480], [640, 480]]]).astype(np.float32)
dists = [np.array([-0.3858, 0.1459, 0.0008, -0.0005, 0.0], np.float32), np.array([-0.28, 0.07, 1e-4, 2e-4], np.float32),
         np.array([0.1, -0.05, 0.002, -0.001, 0.01], np.float32), np.array([-1.9, 3.0, 0.0, 0.0, 0.0], np.float32)]
out = {"K4": K4, "pts": pts, "cv2_version": np.array(cv2.__version__)}
for i, d in enumerate(dists):
    out[f"dist{i}"] = d
    out[f"und{i}"] = cv2.undistortPoints(pts.reshape(-1, 1, 2), K, d, None, K).reshape(-1, 2)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "undistort.npz"), **out)
print("written", {k: getattr(v, "shape", None) for k, v in out.items()})
