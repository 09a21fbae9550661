"""Generates tests/golden/sgbm_golden.npz from the REAL cv2.StereoSGBM (opencv-python-headless).

cv::StereoSGBM is the third-party routine the reference calls for its dense matcher
(src/wass_stereo/wass_stereo.cpp:775-782,837); the reference tree holds no golden vectors for it
(SURVEY.md §4), so these fixtures pin the oracle (oracle/sgbm_oracle.c) and, through it, the CUDA path.
Run:  python tests/golden/make_golden.py      (needs cv2; records cv2.__version__ in the file)
"""
import os
import sys
import numpy as np
import cv2

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from wass_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
KEYS = ("minDisparity", "numDisparities", "blockSize", "P1", "P2", "disp12MaxDiff", "preFilterCap",
        "uniquenessRatio", "speckleWindowSize", "speckleRange", "mode")


def cv2_sgbm(img1, img2, p):
    m = cv2.StereoSGBM_create(p["minDisparity"], p["numDisparities"], p["blockSize"], p["P1"], p["P2"])
    m.setUniquenessRatio(p["uniquenessRatio"])
    m.setDisp12MaxDiff(p["disp12MaxDiff"])
    m.setPreFilterCap(p["preFilterCap"])
    m.setSpeckleRange(p["speckleRange"])
    m.setSpeckleWindowSize(p["speckleWindowSize"])
    m.setMode(cv2.STEREO_SGBM_MODE_HH if p["mode"] == 1 else cv2.STEREO_SGBM_MODE_SGBM)
    return m.compute(img1, img2)


def cases():
    out = []
    # (a) WASS defaults (wass_stereo.cpp:742-761) on the synthetic generator, both modes, WASS padding
    for (W, H, D, seed) in [(96, 64, 16, 1), (160, 72, 64, 2), (200, 48, 128, 3), (330, 40, 256, 4)]:
        r, l, _ = synth.make_pair(W, H, D, seed=seed)
        i1, i2 = synth.pad_for_sgbm(r, l, D)
        for mode in (0, 1):
            p = dict(minDisparity=1, numDisparities=D, blockSize=13, P1=2 * 169, P2=64 * 169, disp12MaxDiff=-1,
                     preFilterCap=60, uniquenessRatio=1, speckleWindowSize=-70, speckleRange=16, mode=mode)
            out.append((i1, i2, p))
    # (b) DISPARITY_OFFSET +/- and negative minDisparity, other windows / penalties
    rng = np.random.default_rng(7)
    for k in range(10):
        H = int(rng.integers(24, 60)); D = 16 * int(rng.integers(1, 6)); minD = int(rng.integers(-3, 5))
        W = int(rng.integers(40, 120))
        r, l, _ = synth.make_pair(W, H, D, seed=100 + k)
        i1, i2 = synth.pad_for_sgbm(r, l, D, disparity_offset=int(rng.integers(-3, 4)))
        win = int(rng.choice([1, 3, 5, 7, 9, 11, 13, 15]))
        P1 = int(rng.integers(1, 400)); P2 = int(rng.integers(P1 + 1, 9000))
        p = dict(minDisparity=minD, numDisparities=D, blockSize=win, P1=P1, P2=P2,
                 disp12MaxDiff=int(rng.integers(-1, 4)), preFilterCap=int(rng.integers(1, 63)),
                 uniquenessRatio=int(rng.integers(0, 20)), speckleWindowSize=0, speckleRange=0, mode=int(k % 2))
        out.append((i1, i2, p))
    # (c) unpadded raw-noise images (worst-case texture), ragged widths
    for k, (H, W, D) in enumerate([(17, 37, 16), (33, 70, 32), (9, 130, 64)]):
        a = rng.integers(60, 200, (H, W)).astype(np.uint8)
        b = np.roll(a, -3, axis=1)
        p = dict(minDisparity=0, numDisparities=D, blockSize=5, P1=8 * 25, P2=32 * 25, disp12MaxDiff=1,
                 preFilterCap=31, uniquenessRatio=10, speckleWindowSize=0, speckleRange=0, mode=k % 2)
        out.append((a, b, p))
    # (d) speckle filter on (cv::filterSpeckles path)
    r, l, _ = synth.make_pair(120, 50, 32, seed=55, noise=6.0)
    i1, i2 = synth.pad_for_sgbm(r, l, 32)
    out.append((i1, i2, dict(minDisparity=1, numDisparities=32, blockSize=5, P1=200, P2=800, disp12MaxDiff=1,
                             preFilterCap=60, uniquenessRatio=5, speckleWindowSize=40, speckleRange=2, mode=0)))
    # (e) images narrower than maxD + blockSize/2 make cv2 raise (stereosgbm.cpp:511) -> no vector; the
    #     boundary returns an error for them (tests/test_capi_symbols.py, test_sgbm_gpu.py)
    return out


def main():
    data = {"cv2_version": np.array(cv2.__version__)}
    cs = cases()
    data["n"] = np.array(len(cs))
    for i, (i1, i2, p) in enumerate(cs):
        data["img1_%d" % i] = i1
        data["img2_%d" % i] = i2
        data["params_%d" % i] = np.array([p[k] for k in KEYS], np.int32)
        data["disp_%d" % i] = cv2_sgbm(i1, i2, p)
    np.savez_compressed(os.path.join(HERE, "sgbm_golden.npz"), **data)
    print("wrote %d cases, cv2 %s" % (len(cs), cv2.__version__))


if __name__ == "__main__":
    main()
