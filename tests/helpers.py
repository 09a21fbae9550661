import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
KEYS = ("minDisparity", "numDisparities", "blockSize", "P1", "P2", "disp12MaxDiff", "preFilterCap",
        "uniquenessRatio", "speckleWindowSize", "speckleRange", "mode")


def load_sgbm_golden():
    z = np.load(os.path.join(GOLDEN, "sgbm_golden.npz"))
    out = []
    for i in range(int(z["n"])):
        p = {k: int(v) for k, v in zip(KEYS, z["params_%d" % i])}
        out.append((z["img1_%d" % i], z["img2_%d" % i], p, z["disp_%d" % i]))
    return out
