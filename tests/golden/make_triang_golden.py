#!/usr/bin/env python
"""Generates tests/golden/triang_golden.npz with oracle/_ref/triang_ref: the reference's OWN
`size_t triangulate( StereoMatchEnv& )` + `StereoMatchEnv::unrectify` (src/wass_stereo/wass_stereo.cpp:299-324, 1039-1386),
cut out of the reference source at build time and compiled with its PovMesh.cpp / triangulate.hpp against the header shim
(oracle/build_ref.sh, oracle/cut_triangulate.awk).  Needs /root/reference and cv2 (for cv::stereoRectify); the .npz travels.

    python tests/golden/make_triang_golden.py
"""
import os
import subprocess
import sys
import tempfile
import numpy as np
import cv2

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.path.join(ROOT, "oracle", "_ref", "triang_ref")


def rig(W, H, rng, tilt):
    f = 1.1 * W
    K0 = np.array([[f, 0, W / 2 + 3.0], [0, f * 1.01, H / 2 - 2.0], [0, 0, 1]])
    K1 = np.array([[f * 0.99, 0, W / 2 - 4.0], [0, f, H / 2 + 1.5], [0, 0, 1]])
    rv = np.array([0.01, -0.02, 0.005]) * tilt
    R, _ = cv2.Rodrigues(rv)
    T = np.array([2.5, 0.03 * tilt, -0.02 * tilt])          # X1 = R X0 + T, cam0 = left (positive x baseline)
    return K0, K1, R, T


def run_case(td, name, W, H, rng, tilt=1.0, cfg="", masks=False, burned=False, custom=False, comp=0.0):
    K0, K1, R, T = rig(W, H, rng, tilt)
    R1, R2, P1, P2, Q, roi1, roi2 = cv2.stereoRectify(K0, np.zeros(5), K1, np.zeros(5), (W, H), R, T, flags=cv2.CALIB_ZERO_DISPARITY, alpha=-1)
    # the common ROI of the reference (wass_stereo.cpp:560-575): same rows, same width
    y0 = max(roi1[1], roi2[1]); y1 = min(roi1[1] + roi1[3], roi2[1] + roi2[3]); w = min(roi1[2], roi2[2])
    roiL = (roi1[0], y0, w, y1 - y0); roiR = (roi2[0], y0, w, y1 - y0)
    left = rng.integers(0, 250, (H, W), dtype=np.uint8)
    right = rng.integers(0, 250, (H, W), dtype=np.uint8)
    if burned:
        left[H // 3:H // 3 + 9, W // 4:W // 2] = 255
        right[H // 2:H // 2 + 5, W // 3:W // 3 + 20] = 255
    lrect = rng.integers(0, 256, (H, W), dtype=np.uint8)
    rrect = rng.integers(0, 256, (H, W), dtype=np.uint8)
    # disparities of a plausible scene (Z between ~8 and ~60 baselines) with holes, sub-pixel values and junk
    disp = (rng.random((H, W)) * 40 + 6).astype(np.float32)
    disp[rng.random((H, W)) < 0.25] = 0
    disp[rng.random((H, W)) < 0.03] = 1.0            # the boundary of "> min_disp"
    disp[rng.random((H, W)) < 0.03] = 900.0          # lands left of the rectified image
    HLi = np.eye(3); HRi = np.eye(3)
    if custom:
        a = 0.02
        HL = np.array([[np.cos(a), -np.sin(a), 2.0], [np.sin(a), np.cos(a), -1.0], [1e-5, -2e-5, 1.0]])
        HR = np.array([[np.cos(a), -np.sin(a), -3.0], [np.sin(a), np.cos(a), -1.0], [-1e-5, 1e-5, 1.0]])
        HLi, HRi = np.linalg.inv(HL), np.linalg.inv(HR)
        cfg += "USE_CUSTOM_STEREORECTIFY=true\n"
    wd = os.path.join(td, name); os.makedirs(os.path.join(wd, "undistorted"))
    lm = rm = None
    if masks:
        lm = (rng.random((H, W)) > 0.2).astype(np.uint8) * 255
        rm = (rng.random((H, W)) > 0.1).astype(np.uint8) * 255
        for fn, m in (("lmask.pgm", lm), ("rmask.pgm", rm)):
            with open(os.path.join(wd, fn), "wb") as f:
                f.write(b"P5 %d %d 255\n" % (W, H)); f.write(m.tobytes())
        cfg += 'LEFT_MASK_IMAGE="lmask.pgm"\nRIGHT_MASK_IMAGE="rmask.pgm"\n'
    camdist = float(np.linalg.norm(T))
    case = os.path.join(wd, "case.bin")
    with open(case, "wb") as f:
        np.array([W, H, W, H] + list(roiL) + list(roiR), np.int32).tofile(f)
        for m in (K0, K1, R, T, R1, R2, P1, P2, HLi, HRi, np.array([comp, camdist])):
            np.ascontiguousarray(m, np.float64).tofile(f)
        for m in (left, right, lrect, rrect):
            m.tofile(f)
        disp.tofile(f)
    cfgf = os.path.join(wd, "cfg.txt"); open(cfgf, "w").write(cfg)
    out = os.path.join(wd, "out.bin")
    subprocess.run([REF, case, cfgf, out], check=True, capture_output=True)
    with open(out, "rb") as f:
        n = int(np.fromfile(f, np.int64, 1)[0]); npts = roiR[2] * roiR[3]
        valid = np.fromfile(f, np.uint8, npts).reshape(roiR[3], roiR[2])
        xyz = np.fromfile(f, np.float64, npts * 3).reshape(roiR[3], roiR[2], 3)
        grey = np.fromfile(f, np.uint8, npts).reshape(roiR[3], roiR[2])
    d = {"K0": K0, "K1": K1, "R": R, "T": T, "R1": R1, "R2": R2, "P1": P1, "P2": P2, "HLi": HLi, "HRi": HRi,
         "roiL": np.array(roiL), "roiR": np.array(roiR), "left": left, "right": right, "disp": disp,
         "scal": np.array([comp, camdist]), "config": np.frombuffer(cfg.encode(), np.uint8), "n": np.array([n]),
         "valid": valid, "xyz": xyz, "grey": grey}
    if masks:
        d["lmask"], d["rmask"] = lm, rm
    print(name, "n =", n, "of", npts)
    return {name + "/" + k: v for k, v in d.items()}


def main():
    subprocess.run(["bash", os.path.join(ROOT, "oracle", "build_ref.sh")], check=True, capture_output=True)
    rng = np.random.default_rng(20261018)
    out, names = {}, []
    with tempfile.TemporaryDirectory() as td:
        cases = [
            ("defaults", dict(W=96, H=72)),
            ("tilted_rig", dict(W=120, H=64, tilt=4.0)),
            ("no_angle_check", dict(W=80, H=60, cfg="TRIANG_MIN_ANGLE=-1\n")),
            ("steep_angle", dict(W=80, H=60, cfg="TRIANG_MIN_ANGLE=1.2\n")),
            ("bbox", dict(W=96, H=72, cfg="TRIANG_BBOX_TOP=10\nTRIANG_BBOX_LEFT=20\nTRIANG_BBOX_BOTTOM=50\nTRIANG_BBOX_RIGHT=70\nTRIANG_MIN_ANGLE=-1\n")),
            ("masks_and_burned", dict(W=96, H=72, masks=True, burned=True, cfg="TRIANG_MIN_ANGLE=-1\n")),
            ("burned_kept", dict(W=96, H=72, burned=True, cfg="DISCARD_BURNED_AREAS=false\nTRIANG_MIN_ANGLE=-1\n")),
            ("scaled_compensated", dict(W=96, H=72, comp=3.0, cfg="DENSE_SCALE=0.75\nTRIANG_MIN_ANGLE=-1\n")),
            ("custom_rectifier", dict(W=96, H=72, custom=True, cfg="TRIANG_MIN_ANGLE=-1\n")),
        ]
        for name, kw in cases:
            out.update(run_case(td, name, rng=rng, **kw)); names.append(name)
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "triang_golden.npz"), **out)
    print("wrote triang_golden.npz")


if __name__ == "__main__":
    sys.exit(main())
