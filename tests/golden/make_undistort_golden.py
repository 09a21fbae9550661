"""Golden vectors for the undistortion of wass_prepare (cv::undistort, wass_prepare.cpp:268), produced with cv2.undistort.
Run in the build container:   python tests/golden/make_undistort_golden.py"""
import os
import numpy as np
import cv2

HERE = os.path.dirname(os.path.abspath(__file__))
rng = np.random.default_rng(21)
cases = {}
k = 0
for (H, W, dist) in [(96, 128, [-0.21, 0.08, 0.001, -0.0007, -0.01]),
                     (75, 211, [0.12, -0.3, 0.0, 0.0]),
                     (64, 5000, [-0.05, 0.01, 0.0005, 0.0003, 0.0]),         # wider than 4096: one-row stripes either way
                     (130, 97, [-0.3, 0.12, 0.002, 0.001, -0.02, 0.01, -0.003, 0.0005]),
                     (40, 60, [])]:
    coarse = rng.integers(0, 256, (H // 5 + 2, W // 5 + 2)).astype(np.float32)
    img = np.clip(cv2.resize(coarse, (W, H), interpolation=cv2.INTER_CUBIC), 0, 255).astype(np.uint8)
    K = np.array([[0.9 * W, 0, W / 2 + 3.3], [0, 0.92 * W, H / 2 - 2.1], [0, 0, 1]])
    d = np.array(dist, np.float64)
    out = cv2.undistort(img, K, d if d.size else None)
    cases["img_%d" % k] = img; cases["K_%d" % k] = K; cases["dist_%d" % k] = d; cases["out_%d" % k] = out
    k += 1
cases["n"] = k
np.savez_compressed(os.path.join(HERE, "undistort_golden.npz"), **cases)
print("cases", k)
