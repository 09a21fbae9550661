#!/usr/bin/env python
"""Generates tests/golden/filters_golden.npz with oracle/_ref/filters_ref: the reference's OWN matrix_dilate_zero,
matrix_erode_zero and clean_and_convert_disparity (src/wass_stereo/wass_stereo.cpp:617-733), cut out of the reference
source at build time and compiled against the header shim (oracle/build_ref.sh).  Needs /root/reference; the .npz travels.

    python tests/golden/make_filters_golden.py
"""
import os
import subprocess
import sys
import tempfile
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.path.join(ROOT, "oracle", "_ref", "filters_ref")


def run(mode, arr, *extra):
    with tempfile.TemporaryDirectory() as td:
        fi, fo = os.path.join(td, "in.raw"), os.path.join(td, "out.raw")
        np.ascontiguousarray(arr).tofile(fi)
        subprocess.run([REF, mode, fi, str(arr.shape[0]), str(arr.shape[1])] + [str(e) for e in extra] + [fo], check=True)
        return np.fromfile(fo, np.float32).reshape(arr.shape)


def main():
    subprocess.run(["bash", os.path.join(ROOT, "oracle", "build_ref.sh")], check=True, capture_output=True)
    rng = np.random.default_rng(20261017)
    out = {}
    names = []
    # float maps with holes: sizes down to the degenerate ones, hole densities from sparse to dominant
    for k, (h, w, pz) in enumerate([(37, 53, 0.1), (64, 64, 0.5), (23, 91, 0.9), (3, 3, 0.4), (2, 7, 0.3), (1, 5, 0.2), (5, 1, 0.2),
                                    (120, 160, 0.02), (40, 40, 0.0), (40, 40, 1.0)]):
        a = (rng.random((h, w)) * 200 + 1).astype(np.float32)
        a[rng.random((h, w)) < pz] = 0
        if k == 2:
            a[rng.random((h, w)) < 0.05] *= -1          # negative entries: "> 0" and "== 0" are different tests
        name = "f%d" % k
        names.append(name)
        out[name + "/src"] = a
        out[name + "/dilate"] = run("dilate", a)
        out[name + "/erode"] = run("erode", a)
        out[name + "/dilate2_erode"] = run("erode", run("dilate", run("dilate", a)))
    cnames = []
    for k, (h, w, mind, nd, off, sc) in enumerate([(31, 47, 1, 256, 0, 1.0), (31, 47, 1, 64, -7, 1.0), (20, 33, -16, 128, 5, 1.0 / 0.75),
                                                   (20, 33, 0, 640, 0, 2.0), (9, 9, 1, 16, 3, 1.0 / 1.5)]):
        d = rng.integers(-17 * 16, (nd + 40) * 16, (h, w)).astype(np.int16)
        d[rng.random((h, w)) < 0.2] = (mind - 1) * 16          # cv2's INVALID marker
        d[0, 0] = mind * 16; d[0, 1] = mind * 16 + 1; d[0, 2] = nd * 16; d[0, 3] = nd * 16 + 1      # the two boundaries
        name = "c%d" % k
        cnames.append(name)
        out[name + "/src"] = d
        out[name + "/args"] = np.array([mind, nd, off, sc], np.float64)
        out[name + "/clean"] = run("clean", d, mind, nd, off, repr(sc))
    out["names"] = np.array(names)
    out["cnames"] = np.array(cnames)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "filters_golden.npz"), **out)
    print("wrote filters_golden.npz:", len(names), "filter cases,", len(cnames), "conversion cases")


if __name__ == "__main__":
    sys.exit(main())
