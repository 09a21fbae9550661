"""ctypes front-end of oracle/sgbm_oracle.c (C restatement of cv::StereoSGBM).

TEST INFRASTRUCTURE ONLY.  Reference call site restated: src/wass_stereo/wass_stereo.cpp:775-782,837.
"""
import ctypes
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

MODE_SGBM = 0
MODE_HH = 1


class Params(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in (
        "minDisparity", "numDisparities", "blockSize", "P1", "P2", "disp12MaxDiff",
        "preFilterCap", "uniquenessRatio", "speckleWindowSize", "speckleRange", "mode")]


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "libwass_oracle.so")
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
        _LIB.sgbm_oracle_compute.restype = ctypes.c_int
        _LIB.sgbm_oracle_compute.argtypes = [
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_size_t,
            ctypes.POINTER(Params), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    return _LIB


def wass_params(num_disp, min_disp=1, win=13, p1_mult=2, p2_mult=64, uniq=1, disp12=-1, cap=60,
                speckle_win=-70, speckle_range=16, mode=MODE_SGBM):
    """SGBM parameters exactly as wass_stereo builds them (wass_stereo.cpp:767-782)."""
    return dict(minDisparity=min_disp, numDisparities=num_disp, blockSize=win,
                P1=p1_mult * win * win, P2=p2_mult * win * win, disp12MaxDiff=disp12,
                preFilterCap=cap, uniquenessRatio=uniq, speckleWindowSize=speckle_win,
                speckleRange=speckle_range, mode=mode)


def compute(img1, img2, params, want_volumes=False, want_raw=False):
    """Returns dict(disp=int16 HxW, maxC=int [, C, S = int16 HxW1xD] [, raw])."""
    img1 = np.ascontiguousarray(img1, dtype=np.uint8)
    img2 = np.ascontiguousarray(img2, dtype=np.uint8)
    assert img1.shape == img2.shape and img1.ndim == 2
    H, W = img1.shape
    p = Params(**params)
    disp = np.empty((H, W), np.int16)
    out = {}
    C = S = raw = None
    if want_volumes:
        minD, D = p.minDisparity, p.numDisparities
        W1 = (W + min(minD, 0)) - max(minD + D, 0)
        C = np.zeros((H, max(W1, 0), D), np.int16)
        S = np.zeros_like(C)
    if want_raw:
        raw = np.empty((H, W), np.int16)
    rc = lib().sgbm_oracle_compute(
        img1.ctypes.data, img2.ctypes.data, H, W, W, ctypes.byref(p), disp.ctypes.data,
        C.ctypes.data if C is not None and C.size else None,
        S.ctypes.data if S is not None and S.size else None,
        raw.ctypes.data if raw is not None else None)
    if rc < 0:
        raise ValueError("sgbm_oracle_compute failed: %d" % rc)
    out["disp"] = disp
    out["maxC"] = rc
    if want_volumes:
        out["C"], out["S"] = C, S
    if want_raw:
        out["raw"] = raw
    return out
