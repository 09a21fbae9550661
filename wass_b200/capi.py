"""ctypes binding of include/wassgpu.h (libwassgpu.so).  Used by tests and bench.py.

No fallback: if the library is missing or no CUDA device is present, the constructors raise.
"""
import ctypes
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libwassgpu.so")

MODE_SGBM = 0
MODE_HH = 1
NUM_STAGES = 8
STAGE_NAMES = ("prefilter", "cost", "aggregate", "wta", "median", "postfilter", "triangulate", "mesh")

ERRORS = {0: "WSG_OK", -1: "WSG_ERR_INVALID_ARG", -2: "WSG_ERR_CUDA", -3: "WSG_ERR_NOMEM",
          -4: "WSG_ERR_TOO_SMALL", -5: "WSG_ERR_STATE"}


class WsgError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("%s (%d): %s" % (ERRORS.get(code, "?"), code, msg))
        self.code = code


class SgbmParams(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in (
        "minDisparity", "numDisparities", "blockSize", "P1", "P2", "disp12MaxDiff",
        "preFilterCap", "uniquenessRatio", "speckleWindowSize", "speckleRange", "mode")]


class SgbmStats(ctypes.Structure):
    _fields_ = [("max_cost", ctypes.c_int), ("out_of_domain", ctypes.c_int), ("kernel_launches", ctypes.c_int),
                ("width1", ctypes.c_int), ("d_padded", ctypes.c_int), ("volume_bytes", ctypes.c_longlong)]


_lib = None


def load():
    """Loads libwassgpu.so; raises if it has not been built (python -m wass_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("libwassgpu.so is not built: run `python -m wass_b200.build` "
                          "(there is no CPU fallback for the wass_stereo hot path)")
    lib = ctypes.CDLL(LIB_PATH)
    vp, ci, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t
    lib.wsg_create.argtypes = [ci, ctypes.POINTER(vp)]
    lib.wsg_destroy.argtypes = [vp]
    lib.wsg_destroy.restype = None
    lib.wsg_set_stream.argtypes = [vp, vp]
    lib.wsg_synchronize.argtypes = [vp]
    lib.wsg_last_error.argtypes = [vp]
    lib.wsg_last_error.restype = ctypes.c_char_p
    lib.wsg_version.restype = ctypes.c_char_p
    lib.wsg_sgbm_compute.argtypes = [vp, vp, vp, ci, ci, sz, ctypes.POINTER(SgbmParams), vp]
    lib.wsg_sgbm_compute_device.argtypes = [vp, vp, vp, ci, ci, sz, ctypes.POINTER(SgbmParams), vp]
    lib.wsg_sgbm_get_stats.argtypes = [vp, ctypes.POINTER(SgbmStats)]
    lib.wsg_sgbm_debug_volumes.argtypes = [vp, vp, vp]
    lib.wsg_profile_enable.argtypes = [vp, ci]
    lib.wsg_profile_reset.argtypes = [vp]
    lib.wsg_profile_get.argtypes = [vp, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ci), ci]
    _lib = lib
    return lib


class Handle:
    """One CUDA device + one stream (wsg_handle)."""

    def __init__(self, device=0):
        self.lib = load()
        h = ctypes.c_void_p()
        rc = self.lib.wsg_create(device, ctypes.byref(h))
        if rc != 0:
            raise WsgError(rc, "wsg_create(device=%d) failed -- a CUDA device is required" % device)
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.wsg_destroy(self.h)
            self.h = None

    __del__ = close

    def _ck(self, rc):
        if rc != 0:
            raise WsgError(rc, self.lib.wsg_last_error(self.h).decode())

    def set_stream(self, cuda_stream_ptr):
        self._ck(self.lib.wsg_set_stream(self.h, ctypes.c_void_p(cuda_stream_ptr)))

    def synchronize(self):
        self._ck(self.lib.wsg_synchronize(self.h))

    # ---- dense matcher ----
    def sgbm_compute(self, img1, img2, params, out=None):
        """Host numpy in, host numpy out (H2D + kernels + D2H). Mirrors StereoSGBM::compute."""
        img1 = np.ascontiguousarray(img1, np.uint8)
        img2 = np.ascontiguousarray(img2, np.uint8)
        if img1.shape != img2.shape or img1.ndim != 2:
            raise ValueError("img1/img2 must be 2-D uint8 arrays of equal shape")
        H, W = img1.shape
        if out is None:
            out = np.empty((H, W), np.int16)
        p = SgbmParams(**params)
        self._ck(self.lib.wsg_sgbm_compute(self.h, img1.ctypes.data, img2.ctypes.data, H, W, W,
                                           ctypes.byref(p), out.ctypes.data))
        return out

    def sgbm_compute_ptr(self, img1_ptr, img2_ptr, rows, cols, stride, params, disp_ptr):
        """Raw host pointers (e.g. pinned torch tensors)."""
        p = SgbmParams(**params)
        self._ck(self.lib.wsg_sgbm_compute(self.h, img1_ptr, img2_ptr, rows, cols, stride, ctypes.byref(p), disp_ptr))

    def sgbm_compute_device(self, d_img1, d_img2, rows, cols, stride, params, d_disp):
        """Device pointers (ints), asynchronous on the handle's stream."""
        p = SgbmParams(**params)
        self._ck(self.lib.wsg_sgbm_compute_device(self.h, d_img1, d_img2, rows, cols, stride, ctypes.byref(p), d_disp))

    def sgbm_stats(self):
        s = SgbmStats()
        self._ck(self.lib.wsg_sgbm_get_stats(self.h, ctypes.byref(s)))
        return {f[0]: getattr(s, f[0]) for f in SgbmStats._fields_}

    def sgbm_debug_volumes(self, rows, w1, D):
        C = np.empty((rows, w1, D), np.int16)
        S = np.empty((rows, w1, D), np.int16)
        self._ck(self.lib.wsg_sgbm_debug_volumes(self.h, C.ctypes.data, S.ctypes.data))
        return C, S

    # ---- profiling ----
    def profile_enable(self, on=True):
        self._ck(self.lib.wsg_profile_enable(self.h, 1 if on else 0))

    def profile_reset(self):
        self._ck(self.lib.wsg_profile_reset(self.h))

    def profile_get(self):
        ms = (ctypes.c_float * NUM_STAGES)()
        ln = (ctypes.c_int * NUM_STAGES)()
        self._ck(self.lib.wsg_profile_get(self.h, ms, ln, NUM_STAGES))
        return {STAGE_NAMES[i]: (float(ms[i]), int(ln[i])) for i in range(NUM_STAGES)}
