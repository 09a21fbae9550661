#!/usr/bin/env python
"""Full-size known answers from the real cv2.StereoSGBM (the routine src/wass_stereo/wass_stereo.cpp:837 calls).

A 2448x2048 or 4096x3000 disparity map is too big for a fixture, so what is committed is its SHA-256 together with
the SHA-256 of the two input images: tests/test_fullsize_parity.py regenerates the seeded frame, and if its input
hash matches (same numpy / cv2 resize arithmetic on the box) the GPU disparity must hash to the cv2 value recorded
here; if the inputs differ on that machine it runs cv2 there instead.  The benchmark frame of bench.py (seed 0) is
among the cases, so the timed frame itself is pinned.

    python tests/golden/make_fullsize_hashes.py        # ~4 min, ~30 GB of RAM for the 4096x3000x512 MODE_HH case
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

# (name, W, H, D, mode, seed)  -- mode 1 = MODE_HH (8 paths), 0 = MODE_SGBM (5 paths, the reference's default)
CASES = [
    ("bench_frame_seed0_hh", 2448, 2048, 256, 1, 0),
    ("bench_frame_seed1_hh", 2448, 2048, 256, 1, 1),
    ("bench_frame_seed2_hh", 2448, 2048, 256, 1, 2),
    ("config2_hh", 2448, 2048, 256, 1, 7),
    ("config2_sgbm", 2448, 2048, 256, 0, 7),
    ("config4_hh", 4096, 3000, 512, 1, 7),
]


def sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(str(a.shape).encode() + str(a.dtype).encode())
        h.update(a.tobytes())
    return h.hexdigest()


def wass_params(D, mode):
    win = 13
    return dict(minDisparity=1, numDisparities=D, blockSize=win, P1=2 * win * win, P2=64 * win * win, disp12MaxDiff=-1,
                preFilterCap=60, uniquenessRatio=1, speckleWindowSize=-70, speckleRange=16, mode=mode)


def cv2_sgbm(i1, i2, p):
    import cv2
    cv2.setNumThreads(1)
    m = cv2.StereoSGBM_create(p["minDisparity"], p["numDisparities"], p["blockSize"], p["P1"], p["P2"])
    m.setUniquenessRatio(p["uniquenessRatio"]); m.setDisp12MaxDiff(p["disp12MaxDiff"])
    m.setPreFilterCap(p["preFilterCap"]); m.setSpeckleRange(p["speckleRange"])
    m.setSpeckleWindowSize(p["speckleWindowSize"])
    m.setMode(cv2.STEREO_SGBM_MODE_HH if p["mode"] == 1 else cv2.STEREO_SGBM_MODE_SGBM)
    return m.compute(i1, i2)


def main():
    import cv2
    from wass_b200 import synth
    out = {"cv2": cv2.__version__, "numpy": np.__version__, "cases": {}}
    for name, W, H, D, mode, seed in CASES:
        r, l, _ = synth.make_pair(W, H, D, seed=seed)
        i1, i2 = synth.pad_for_sgbm(r, l, D)
        t = time.time()
        disp = cv2_sgbm(i1, i2, wass_params(D, mode))
        out["cases"][name] = {"W": W, "H": H, "D": D, "mode": mode, "seed": seed, "inputs_sha256": sha(i1, i2),
                              "disp_sha256": sha(disp), "valid_fraction": float((disp[:, D:] > 16).mean()),
                              "cv2_seconds": round(time.time() - t, 1)}
        print(name, out["cases"][name], flush=True)
    with open(os.path.join(ROOT, "tests", "golden", "fullsize_hashes.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
