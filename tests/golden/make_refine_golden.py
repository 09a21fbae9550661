"""Golden vectors for the optional disparity refinement (wass_stereo.cpp:941-986), produced with the OpenCV routines the
reference calls: cv2.medianBlur, cv2.Sobel, cv2.connectedComponentsWithStats.  Run in the build container:

    python tests/golden/make_refine_golden.py
"""
import os
import numpy as np
import cv2

HERE = os.path.dirname(os.path.abspath(__file__))


def reference_refine(disp, median_wsize, thr):
    d = disp.copy()
    if median_wsize >= 3:
        d = cv2.medianBlur(d, median_wsize)
    if thr > 0:
        gx = cv2.Sobel(d, cv2.CV_32F, 1, 0)
        gy = cv2.Sobel(d, cv2.CV_32F, 0, 1)
        g2 = gx * gx + gy * gy
        d[g2 > thr] = 0.0
        mask = (d != 0).astype(np.uint8)
        n, labels, stats, _ = cv2.connectedComponentsWithStats(mask)
        max_area, best = 0, None
        for i in range(1, n):
            if stats[i, cv2.CC_STAT_AREA] > max_area:
                max_area, best = stats[i, cv2.CC_STAT_AREA], i
        if best is not None:
            d[labels != best] = 0.0
        # (no component at all: the reference's `1 - biggestcomp_mask` on an empty Mat throws; unreachable with data)
    return d


rng = np.random.default_rng(11)
cases = {}
k = 0
for (H, W) in [(48, 64), (37, 91), (64, 64)]:
    y, x = np.mgrid[0:H, 0:W].astype(np.float32)
    base = (20 + 0.8 * y + 3 * np.sin(x / 5.0)).astype(np.float32)
    base += rng.normal(0, 0.3, base.shape).astype(np.float32)
    base[rng.random(base.shape) < 0.12] = 0          # holes
    base[:, W // 2 - 1:W // 2 + 1] = 0                # a gap that splits the map in two components
    base[H // 3, : W // 2 - 1] += 40                  # a step edge: large gradient
    for (m, t) in [(0, 0), (3, 0), (5, 0), (0, 50), (3, 200), (5, 20)]:
        cases["in_%d" % k] = base
        cases["par_%d" % k] = np.array([m, t])
        cases["out_%d" % k] = reference_refine(base, m, t)
        k += 1
# a tie: two components of equal area; cv2 keeps the one with the smaller label
tie = np.zeros((12, 20), np.float32)
tie[2:5, 12:16] = 7.0
tie[6:10, 2:5] = 9.0
assert (tie[2:5, 12:16] != 0).sum() == (tie[6:10, 2:5] != 0).sum()
cases["in_%d" % k] = tie; cases["par_%d" % k] = np.array([0, 1000000]); cases["out_%d" % k] = reference_refine(tie, 0, 1000000); k += 1
cases["n"] = k
np.savez_compressed(os.path.join(HERE, "refine_golden.npz"), **cases)
print("cases", k)
