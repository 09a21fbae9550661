"""ctypes binding of include/wassgpu.h (libwassgpu.so).  Used by tests and bench.py.

No fallback: if the library is missing or no CUDA device is present, the constructors raise.
"""
import ctypes
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("WSG_LIB") or os.path.join(_HERE, "libwassgpu.so")     # WSG_LIB: an experiment build (tools/)

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
                ("width1", ctypes.c_int), ("d_padded", ctypes.c_int), ("agg_impl", ctypes.c_int),
                ("volume_bytes", ctypes.c_longlong)]


class DenseParams(ctypes.Structure):
    _fields_ = [("MIN_DISPARITY", ctypes.c_int), ("MAX_DISPARITY", ctypes.c_int), ("WINSIZE", ctypes.c_int),
                ("DENSE_SCALE", ctypes.c_double), ("DISPARITY_OFFSET", ctypes.c_int), ("DISP_DILATE_STEPS", ctypes.c_int),
                ("DISP_EROSION_STEPS", ctypes.c_int), ("DENSE_P1_MULT", ctypes.c_int), ("DENSE_P2_MULT", ctypes.c_int),
                ("DENSE_UNIQUENESS_RATIO", ctypes.c_int), ("DENSE_DISP12MAXDIFF", ctypes.c_int),
                ("DENSE_PREFILTER_CAP", ctypes.c_int), ("DENSE_SPECKLE_RANGE", ctypes.c_int),
                ("DENSE_SPECKLE_WINDOW_SIZE", ctypes.c_int), ("mode", ctypes.c_int),
                ("MEDIAN_FILTER_WSIZE", ctypes.c_int), ("DENSE_DISPARITY_BIGGEST_COMPONENT_THRESHOLD", ctypes.c_int)]


class Calib(ctypes.Structure):
    _fields_ = [("K0", ctypes.c_double * 9), ("K1", ctypes.c_double * 9), ("R", ctypes.c_double * 9), ("T", ctypes.c_double * 3),
                ("R1", ctypes.c_double * 9), ("R2", ctypes.c_double * 9), ("P1", ctypes.c_double * 12), ("P2", ctypes.c_double * 12),
                ("roi_left", ctypes.c_int * 4), ("roi_right", ctypes.c_int * 4),
                ("left_cols", ctypes.c_int), ("left_rows", ctypes.c_int), ("right_cols", ctypes.c_int), ("right_rows", ctypes.c_int),
                ("rect_cols", ctypes.c_int), ("rect_rows", ctypes.c_int),
                ("use_homographies", ctypes.c_int), ("HLi", ctypes.c_double * 9), ("HRi", ctypes.c_double * 9)]


class TriParams(ctypes.Structure):
    _fields_ = [("TRIANG_MIN_ANGLE", ctypes.c_double), ("TRIANG_BBOX_TOP", ctypes.c_double), ("TRIANG_BBOX_LEFT", ctypes.c_double),
                ("TRIANG_BBOX_RIGHT", ctypes.c_double), ("TRIANG_BBOX_BOTTOM", ctypes.c_double),
                ("DISCARD_BURNED_AREAS", ctypes.c_int), ("disparity_compensation", ctypes.c_int),
                ("DENSE_SCALE", ctypes.c_double), ("cam_distance", ctypes.c_double)]


class RefineParams(ctypes.Structure):
    _fields_ = [("PLANE_REFINE_XMIN", ctypes.c_double), ("PLANE_REFINE_XMAX", ctypes.c_double),
                ("PLANE_REFINE_YMIN", ctypes.c_double), ("PLANE_REFINE_YMAX", ctypes.c_double),
                ("PLANE_REFINEMENT_MAX_DISTANCE", ctypes.c_double), ("PLANE_WEIGHT_PROPORTIONAL_TO_DISTANCE", ctypes.c_int),
                ("PLANE_USE_CENTRAL_THIRD_ONLY", ctypes.c_int)]


def make_calib(c, left_shape, right_shape, rect_shape):
    """dict(K0,K1,R,T,R1,R2,P1,P2,roi_left,roi_right) -> Calib"""
    k = Calib()
    custom = "HLi" in c                      # USE_CUSTOM_STEREORECTIFY: homographies instead of R1,R2,P1,P2
    k.use_homographies = int(custom)
    names = (("K0", 9), ("K1", 9), ("R", 9), ("T", 3)) + ((("HLi", 9), ("HRi", 9)) if custom else
                                                             (("R1", 9), ("R2", 9), ("P1", 12), ("P2", 12)))
    for name, n in names:
        arr = np.asarray(c[name], np.float64).reshape(-1)
        assert arr.size == n, name
        getattr(k, name)[:] = list(arr)
    k.roi_left[:] = [int(v) for v in c["roi_left"]]
    k.roi_right[:] = [int(v) for v in c["roi_right"]]
    k.left_rows, k.left_cols = left_shape
    k.right_rows, k.right_cols = right_shape
    k.rect_rows, k.rect_cols = rect_shape
    return k


AGG_PER_DIRECTION, AGG_SWEEPS, AGG_SWEEPS_WTA = 0, 1, 2
MAX_BATCH = 64

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
    lib.wsg_sgbm_compute_batch.argtypes = [vp, ci, ctypes.POINTER(vp), ctypes.POINTER(vp), ci, ci, sz, ctypes.POINTER(SgbmParams),
                                           ctypes.POINTER(vp)]
    lib.wsg_sgbm_compute_batch_device.argtypes = [vp, ci, vp, vp, sz, ci, ci, sz, ctypes.POINTER(SgbmParams), vp]
    lib.wsg_sgbm_batch_submit.argtypes = [vp, ci, ci, ctypes.POINTER(vp), ctypes.POINTER(vp), ci, ci, sz, ctypes.POINTER(SgbmParams),
                                          ctypes.POINTER(vp)]
    lib.wsg_sgbm_batch_wait.argtypes = [vp, ci]
    lib.wsg_sgbm_get_stats.argtypes = [vp, ctypes.POINTER(SgbmStats)]
    lib.wsg_sgbm_debug_volumes.argtypes = [vp, vp, vp]
    lib.wsg_sgbm_set_sweep_workers.argtypes = [vp, ci]
    lib.wsg_sgbm_set_impl.argtypes = [vp, ci]
    lib.wsg_profile_enable.argtypes = [vp, ci]
    lib.wsg_profile_reset.argtypes = [vp]
    lib.wsg_profile_get.argtypes = [vp, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ci), ci]
    u64p = ctypes.POINTER(ctypes.c_ulonglong)
    dp = ctypes.POINTER(ctypes.c_double)
    lib.wsg_dense_params_default.argtypes = [ctypes.POINTER(DenseParams)]
    lib.wsg_dense_params_default.restype = None
    lib.wsg_tri_params_default.argtypes = [ctypes.POINTER(TriParams)]
    lib.wsg_tri_params_default.restype = None
    lib.wsg_refine_params_default.argtypes = [ctypes.POINTER(RefineParams)]
    lib.wsg_refine_params_default.restype = None
    lib.wsg_dense_stereo.argtypes = [vp, vp, vp, ci, ci, sz, ctypes.POINTER(DenseParams), vp, vp]
    lib.wsg_dense_stereo_batch.argtypes = [vp, ci, ctypes.POINTER(vp), ctypes.POINTER(vp), ci, ci, sz, ctypes.POINTER(DenseParams),
                                           ctypes.POINTER(vp)]
    lib.wsg_dense_select.argtypes = [vp, ci]
    lib.wsg_disparity_postprocess.argtypes = [vp, vp, ci, ci, ci, ci, ci, ctypes.c_double, ci, ci, vp]
    lib.wsg_disparity_postprocess_resized.argtypes = [vp, vp, ci, ci, ci, ci, ci, ctypes.c_double, ci, ci, vp, ci, ci]
    lib.wsg_dense_scaled_size.argtypes = [ci, ci, ctypes.c_double, ctypes.POINTER(ci), ctypes.POINTER(ci)]
    lib.wsg_dense_scaled_size.restype = None
    lib.wsg_resize_u8_cubic.argtypes = [vp, vp, ci, ci, sz, ctypes.c_double, ctypes.c_double, vp, ci, ci]
    lib.wsg_resize_f32.argtypes = [vp, vp, ci, ci, vp, ci, ci, ci]
    lib.wsg_disparity_refine.argtypes = [vp, vp, ci, ci, ci, ci]
    lib.wsg_triangulate.argtypes = [vp, vp, vp, vp, vp, vp, ctypes.POINTER(Calib), ctypes.POINTER(TriParams), u64p]
    lib.wsg_triangulate_from_dense.argtypes = [vp, vp, vp, vp, vp, ctypes.POINTER(Calib), ctypes.POINTER(TriParams), u64p]
    lib.wsg_mesh_upload.argtypes = [vp, ci, ci, vp, vp, vp]
    lib.wsg_mesh_download.argtypes = [vp, vp, vp, vp]
    lib.wsg_mesh_size.argtypes = [vp, ctypes.POINTER(ci), ctypes.POINTER(ci), u64p]
    lib.wsg_mesh_zgap_percentile.argtypes = [vp, ctypes.c_double, dp]
    lib.wsg_mesh_biggest_component.argtypes = [vp, ctypes.c_double, u64p]
    lib.wsg_mesh_ransac_plane.argtypes = [vp, vp, ci, ctypes.c_double, dp, ctypes.POINTER(ci), u64p]
    lib.wsg_ransac_draw.argtypes = [ci, ci, ci, vp]
    lib.wsg_mesh_crop_plane.argtypes = [vp, dp, ctypes.c_double, u64p]
    lib.wsg_mesh_refine_plane.argtypes = [vp, ctypes.POINTER(RefineParams), dp, u64p]
    lib.wsg_mesh_refine_inliers.argtypes = [vp, ctypes.POINTER(RefineParams), ci, vp, sz, u64p, u64p]
    lib.wsg_rt_from_plane.argtypes = [dp, dp, dp, dp, dp]
    lib.wsg_rt_from_plane.restype = None
    lib.wsg_mesh_export_xyzc.argtypes = [vp, dp, vp, sz, ctypes.POINTER(sz)]
    lib.wsg_mesh_export_xyzbin.argtypes = [vp, vp, sz, ctypes.POINTER(sz)]
    lib.wsg_xyzc_decode_align.argtypes = [vp, vp, sz, dp, ctypes.c_double, vp, sz, ctypes.POINTER(sz)]
    lib.wsg_mesh_aligned_points.argtypes = [vp, dp, dp, ctypes.c_double, vp, sz, ctypes.POINTER(sz)]
    ip = ctypes.POINTER(ci)
    lib.wsg_stereo_rectify.argtypes = [dp, dp, dp, dp, ci, ci, dp, dp, dp, dp, ip, ip]
    lib.wsg_rectify_image.argtypes = [vp, vp, ci, ci, sz, dp, dp, dp, vp]
    lib.wsg_stereo_rectify_custom.argtypes = [dp, dp, dp, dp, ctypes.c_double, ci, ci, dp, dp, ip, dp]
    lib.wsg_warp_perspective.argtypes = [vp, vp, ci, ci, sz, dp, vp]
    lib.wsg_undistort_image.argtypes = [vp, vp, ci, ci, sz, dp, dp, ci, vp]
    lib.wsg_clahe_image.argtypes = [vp, vp, ci, ci, sz, ctypes.c_double, ci, vp]
    lib.wsg_prepare_image.argtypes = [vp, vp, ci, ci, sz, ci, ctypes.c_double, dp, dp, ci, vp]
    lib.wsg_plane_mean_accumulate.argtypes = [dp, dp]
    lib.wsg_plane_mean_accumulate.restype = None
    lib.wsg_plane_mean_finish.argtypes = [dp, dp]
    lib.wsg_plane_mean_finish.restype = None
    lib.wsg_nccl_unique_id.argtypes = [ctypes.c_char_p]
    lib.wsg_nccl_comm_create.argtypes = [ci, ci, ci, ctypes.c_char_p, ctypes.POINTER(vp)]
    lib.wsg_nccl_comm_destroy.argtypes = [vp]
    lib.wsg_nccl_comm_destroy.restype = None
    lib.wsg_plane_allreduce.argtypes = [vp, vp, dp, dp, ctypes.POINTER(ctypes.c_longlong)]
    lib.wsg_plane_allgather.argtypes = [vp, vp, ci, dp, ci, dp]
    lib.wsg_collective_last_error.restype = ctypes.c_char_p
    _lib = lib
    return lib


def dense_params(**kw):
    p = DenseParams()
    load().wsg_dense_params_default(ctypes.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def tri_params(**kw):
    p = TriParams()
    load().wsg_tri_params_default(ctypes.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def refine_params(**kw):
    p = RefineParams()
    load().wsg_refine_params_default(ctypes.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def ransac_draw(width, height, rounds):
    """libc rand()-driven triple draw (PovMesh.cpp:678-692); seed with libc srand first."""
    t = np.zeros((rounds, 6), np.int32)
    rc = load().wsg_ransac_draw(width, height, rounds, t.ctypes.data)
    if rc:
        raise WsgError(rc, "wsg_ransac_draw")
    return t


def _darr(a, n):
    a = np.asarray(a, np.float64).reshape(-1)
    assert a.size == n
    return (ctypes.c_double * n)(*a)


def stereo_rectify(K0, K1, R, T, width, height):
    """cv::stereoRectify as wass_stereo calls it (flags=0, alpha=1, zero distortion). Host only."""
    R1, R2 = (ctypes.c_double * 9)(), (ctypes.c_double * 9)()
    P1, P2 = (ctypes.c_double * 12)(), (ctypes.c_double * 12)()
    r1, r2 = (ctypes.c_int * 4)(), (ctypes.c_int * 4)()
    rc = load().wsg_stereo_rectify(_darr(K0, 9), _darr(K1, 9), _darr(R, 9), _darr(T, 3), width, height, R1, R2, P1, P2, r1, r2)
    if rc:
        raise WsgError(rc, "wsg_stereo_rectify")
    return dict(R1=np.array(list(R1)).reshape(3, 3), R2=np.array(list(R2)).reshape(3, 3), P1=np.array(list(P1)).reshape(3, 4),
                P2=np.array(list(P2)).reshape(3, 4), roi1=tuple(r1), roi2=tuple(r2))


def stereo_rectify_custom(K0, K1, R, T, width, height, rot_angle=0.0):
    """stereoRectifyUndistorted (src/wass_stereo/stereorectify.cpp:57-244); R,T take camera-1 points into camera 0
    (wass_stereo.cpp:502 passes Rinv, Tinv).  Host only.  Returns dict(H0, H1, roi, angle)."""
    H0, H1 = (ctypes.c_double * 9)(), (ctypes.c_double * 9)()
    roi = (ctypes.c_int * 4)()
    ang = ctypes.c_double()
    rc = load().wsg_stereo_rectify_custom(_darr(K0, 9), _darr(K1, 9), _darr(R, 9), _darr(T, 3), float(rot_angle), width, height,
                                          H0, H1, roi, ctypes.cast(ctypes.byref(ang), ctypes.POINTER(ctypes.c_double)))
    if rc:
        raise WsgError(rc, "wsg_stereo_rectify_custom")
    return dict(H0=np.array(list(H0)).reshape(3, 3), H1=np.array(list(H1)).reshape(3, 3), roi=tuple(roi), angle=ang.value)


def rt_from_plane(plane):
    R, T, Ri, Ti = (ctypes.c_double * 9)(), (ctypes.c_double * 3)(), (ctypes.c_double * 9)(), (ctypes.c_double * 3)()
    load().wsg_rt_from_plane(_d4(plane), R, T, Ri, Ti)
    return (np.array(list(R)).reshape(3, 3), np.array(list(T)), np.array(list(Ri)).reshape(3, 3), np.array(list(Ti)))


def plane_mean(planes):
    acc = (ctypes.c_double * 5)(0, 0, 0, 0, 0)
    for p in planes:
        load().wsg_plane_mean_accumulate(acc, _d4(p))
    mean = (ctypes.c_double * 4)()
    load().wsg_plane_mean_finish(acc, mean)
    return np.array(list(mean)), np.array(list(acc))


def _d4(a):
    return (ctypes.c_double * 4)(*[float(v) for v in a])


NCCL_UNIQUE_ID_BYTES = 128


def nccl_unique_id():
    """128 bytes from ncclGetUniqueId (rank 0 calls this and hands them to the other ranks)."""
    buf = ctypes.create_string_buffer(NCCL_UNIQUE_ID_BYTES)
    rc = load().wsg_nccl_unique_id(buf)
    if rc != 0:
        raise WsgError(rc, load().wsg_collective_last_error().decode())
    return buf.raw


class NcclComm:
    """An NCCL communicator owned through the C ABI (wsg_nccl_comm_create): one rank per GPU."""

    def __init__(self, device, nranks, rank, unique_id):
        self.lib = load()
        c = ctypes.c_void_p()
        rc = self.lib.wsg_nccl_comm_create(device, nranks, rank, bytes(unique_id), ctypes.byref(c))
        if rc != 0:
            raise WsgError(rc, self.lib.wsg_collective_last_error().decode())
        self.c, self.nranks, self.rank = c, nranks, rank

    def close(self):
        if getattr(self, "c", None):
            self.lib.wsg_nccl_comm_destroy(self.c)
            self.c = None


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

    def sgbm_compute_batch(self, imgs1, imgs2, params):
        """n frames of equal shape in one call (wsg_sgbm_compute_batch): lists of host arrays in, list of disparities out."""
        a = [np.ascontiguousarray(i, np.uint8) for i in imgs1]
        b = [np.ascontiguousarray(i, np.uint8) for i in imgs2]
        n = len(a)
        if n == 0 or len(b) != n or any(x.shape != a[0].shape or x.ndim != 2 for x in a + b):
            raise ValueError("imgs1/imgs2 must be equally long lists of 2-D uint8 arrays of one shape")
        H, W = a[0].shape
        out = [np.empty((H, W), np.int16) for _ in range(n)]
        arr = ctypes.c_void_p * n
        p = SgbmParams(**params)
        self._ck(self.lib.wsg_sgbm_compute_batch(self.h, n, arr(*[x.ctypes.data for x in a]), arr(*[x.ctypes.data for x in b]),
                                                 H, W, W, ctypes.byref(p), arr(*[x.ctypes.data for x in out])))
        return out

    def sgbm_compute_batch_ptr(self, n, img1_ptrs, img2_ptrs, rows, cols, stride, params, disp_ptrs):
        """Raw host pointers (e.g. pinned torch tensors), n of each."""
        arr = ctypes.c_void_p * n
        p = SgbmParams(**params)
        self._ck(self.lib.wsg_sgbm_compute_batch(self.h, n, arr(*img1_ptrs), arr(*img2_ptrs), rows, cols, stride, ctypes.byref(p),
                                                 arr(*disp_ptrs)))

    def sgbm_batch_submit(self, slot, n, img1_ptrs, img2_ptrs, rows, cols, stride, params, disp_ptrs):
        """Asynchronous: enqueue H2D + matcher + D2H of one batch (raw host pointers, pinned) and return; slot 0 or 1."""
        arr = ctypes.c_void_p * n
        p = SgbmParams(**params)
        self._ck(self.lib.wsg_sgbm_batch_submit(self.h, slot, n, arr(*img1_ptrs), arr(*img2_ptrs), rows, cols, stride, ctypes.byref(p),
                                                arr(*disp_ptrs)))

    def sgbm_batch_wait(self, slot):
        self._ck(self.lib.wsg_sgbm_batch_wait(self.h, slot))

    def sgbm_compute_batch_device(self, n, d_img1, d_img2, frame_stride, rows, cols, stride, params, d_disp):
        """Device pointers (ints) to n frames `frame_stride` bytes apart; d_disp: n x rows x cols int16.  Asynchronous."""
        p = SgbmParams(**params)
        self._ck(self.lib.wsg_sgbm_compute_batch_device(self.h, n, d_img1, d_img2, frame_stride, rows, cols, stride,
                                                        ctypes.byref(p), d_disp))

    # ---- the cross-frame reduction (NCCL, on this handle's stream) ----
    def plane_allreduce(self, comm, acc):
        """acc: this rank's 5 NaN-aware sums (plane_mean(...)[1]).  Returns (mean plane[4], frames in it) on every rank."""
        a = (ctypes.c_double * 5)(*[float(v) for v in acc])
        mean, n = (ctypes.c_double * 4)(), ctypes.c_longlong()
        self._ck(self.lib.wsg_plane_allreduce(self.h, comm.c, a, mean, ctypes.byref(n)))
        return np.array(list(mean)), int(n.value)

    def plane_allgather(self, comm, planes):
        """planes: n_local x 4 on every rank (same n_local).  Returns nranks x n_local x 4, rank order."""
        pl = np.ascontiguousarray(planes, np.float64).reshape(-1, 4)
        out = np.empty((comm.nranks, pl.shape[0], 4), np.float64)
        dp = ctypes.POINTER(ctypes.c_double)
        self._ck(self.lib.wsg_plane_allgather(self.h, comm.c, comm.nranks, pl.ctypes.data_as(dp), pl.shape[0], out.ctypes.data_as(dp)))
        return out

    def sgbm_stats(self):
        s = SgbmStats()
        self._ck(self.lib.wsg_sgbm_get_stats(self.h, ctypes.byref(s)))
        return {f[0]: getattr(s, f[0]) for f in SgbmStats._fields_}

    def sgbm_debug_volumes(self, rows, w1, D, want_S=True):
        C = np.empty((rows, w1, D), np.int16)
        S = np.empty((rows, w1, D), np.int16) if want_S else None
        self._ck(self.lib.wsg_sgbm_debug_volumes(self.h, C.ctypes.data, S.ctypes.data if want_S else None))
        return C, S

    def sgbm_set_sweep_workers(self, max_sms):
        """Cap on the SMs the fused sweeps occupy (0 = all)."""
        self._ck(self.lib.wsg_sgbm_set_sweep_workers(self.h, int(max_sms)))

    def sgbm_set_impl(self, impl):
        """AGG_PER_DIRECTION | AGG_SWEEPS | AGG_SWEEPS_WTA (default): same results, different HBM traffic."""
        self._ck(self.lib.wsg_sgbm_set_impl(self.h, int(impl)))

    # ---- dense stage ----
    def dense_stereo(self, left_crop, right_crop, params, want_disp16=False, want_host=True):
        """Mirrors sgbm_dense_stereo (wass_stereo.cpp:764-1020): crops in, float ROI disparity out.
        want_host=False keeps the disparity on the device only (for triangulate_from_dense) and returns None."""
        left_crop = np.ascontiguousarray(left_crop, np.uint8)
        right_crop = np.ascontiguousarray(right_crop, np.uint8)
        H, W = left_crop.shape
        out = np.empty((H, W), np.float32) if want_host else None
        hs, ws = ctypes.c_int(), ctypes.c_int()      # the matcher runs on the DENSE_SCALE-resized crops
        self.lib.wsg_dense_scaled_size(H, W, params.DENSE_SCALE, ctypes.byref(hs), ctypes.byref(ws))
        d16 = np.empty((hs.value, ws.value), np.int16) if want_disp16 else None
        self._ck(self.lib.wsg_dense_stereo(self.h, left_crop.ctypes.data, right_crop.ctypes.data, H, W, W, ctypes.byref(params),
                                           out.ctypes.data if want_host else None, d16.ctypes.data if want_disp16 else None))
        return (out, d16) if want_disp16 else out

    def dense_stereo_batch(self, left_crops, right_crops, params, want_host=False):
        """n pairs of crops of one size through ONE batched matcher run (wsg_dense_stereo_batch).  The float ROI disparities
        stay on the device: dense_select(f) makes frame f the input of triangulate_from_dense.  want_host: also return them."""
        a = [np.ascontiguousarray(c, np.uint8) for c in left_crops]
        b = [np.ascontiguousarray(c, np.uint8) for c in right_crops]
        n = len(a)
        H, W = a[0].shape
        arr = ctypes.c_void_p * n
        outs = [np.empty((H, W), np.float32) for _ in range(n)] if want_host else None
        self._ck(self.lib.wsg_dense_stereo_batch(self.h, n, arr(*[x.ctypes.data for x in a]), arr(*[x.ctypes.data for x in b]), H, W, W,
                                                 ctypes.byref(params), arr(*[x.ctypes.data for x in outs]) if want_host else None))
        return outs

    def dense_select(self, frame):
        self._ck(self.lib.wsg_dense_select(self.h, int(frame)))

    def disparity_postprocess(self, disp16_roi, min_disp, num_disp, disparity_offset=0, dense_scale=1.0, dilate=1, erode=2,
                              out_size=None):
        """wass_stereo.cpp:853-928; out_size=(rows, cols) of the ROI when dense_scale != 1."""
        d = np.ascontiguousarray(disp16_roi, np.int16)
        H, W = d.shape
        oh, ow = (H, W) if out_size is None else out_size
        out = np.empty((oh, ow), np.float32)
        self._ck(self.lib.wsg_disparity_postprocess_resized(self.h, d.ctypes.data, H, W, min_disp, num_disp, disparity_offset,
                                                            dense_scale, dilate, erode, out.ctypes.data, oh, ow))
        return out

    def resize_u8_cubic(self, img, fx, fy):
        """cv::resize(img, None, fx=fx, fy=fy, interpolation=INTER_CUBIC) on uint8 (wass_stereo.cpp:790-795)."""
        a = np.ascontiguousarray(img, np.uint8)
        H, W = a.shape
        dh, dw = int(np.rint(H * fy)), int(np.rint(W * fx))
        out = np.empty((dh, dw), np.uint8)
        self._ck(self.lib.wsg_resize_u8_cubic(self.h, a.ctypes.data, H, W, W, fx, fy, out.ctypes.data, dh, dw))
        return out

    def resize_f32(self, img, dst_rows, dst_cols, interpolation="cubic"):
        """cv::resize(img, (dst_cols, dst_rows), interpolation=INTER_CUBIC|INTER_NEAREST) on float32 (:903-904)."""
        a = np.ascontiguousarray(img, np.float32)
        out = np.empty((dst_rows, dst_cols), np.float32)
        self._ck(self.lib.wsg_resize_f32(self.h, a.ctypes.data, a.shape[0], a.shape[1], out.ctypes.data, dst_rows, dst_cols,
                                         {"nearest": 0, "cubic": 2}[interpolation]))
        return out

    # ---- triangulation + mesh ----
    def disparity_refine(self, disp_roi, median_wsize=0, bc_threshold=0):
        """wass_stereo.cpp:941-986: optional float median (3|5) and gradient mask + biggest 8-connected component."""
        d = np.ascontiguousarray(disp_roi, np.float32).copy()
        self._ck(self.lib.wsg_disparity_refine(self.h, d.ctypes.data, d.shape[0], d.shape[1], int(median_wsize), int(bc_threshold)))
        return d

    def triangulate(self, disparity, left, right, calib, params=None, left_mask=None, right_mask=None):
        disparity = np.ascontiguousarray(disparity, np.float32)
        left = np.ascontiguousarray(left, np.uint8)
        right = np.ascontiguousarray(right, np.uint8)
        c = make_calib(calib, left.shape, right.shape, disparity.shape) if isinstance(calib, dict) else calib
        p = params or tri_params()
        n = ctypes.c_ulonglong()
        lm = np.ascontiguousarray(left_mask, np.uint8) if left_mask is not None else None
        rm = np.ascontiguousarray(right_mask, np.uint8) if right_mask is not None else None
        self._ck(self.lib.wsg_triangulate(self.h, disparity.ctypes.data, left.ctypes.data, right.ctypes.data,
                                          lm.ctypes.data if lm is not None else None, rm.ctypes.data if rm is not None else None,
                                          ctypes.byref(c), ctypes.byref(p), ctypes.byref(n)))
        return n.value

    def triangulate_from_dense(self, left, right, calib, rect_shape, params=None):
        left = np.ascontiguousarray(left, np.uint8)
        right = np.ascontiguousarray(right, np.uint8)
        c = make_calib(calib, left.shape, right.shape, rect_shape) if isinstance(calib, dict) else calib
        p = params or tri_params()
        n = ctypes.c_ulonglong()
        self._ck(self.lib.wsg_triangulate_from_dense(self.h, left.ctypes.data, right.ctypes.data, None, None,
                                                     ctypes.byref(c), ctypes.byref(p), ctypes.byref(n)))
        return n.value

    def mesh_upload(self, valid, p3d, grey=None):
        valid = np.ascontiguousarray(valid, np.uint8)
        p3d = np.ascontiguousarray(p3d, np.float64)
        H, W = valid.shape
        g = np.ascontiguousarray(grey, np.uint8) if grey is not None else None
        self._ck(self.lib.wsg_mesh_upload(self.h, W, H, valid.ctypes.data, p3d.ctypes.data, g.ctypes.data if g is not None else None))

    def mesh_size(self):
        w, hh, n = ctypes.c_int(), ctypes.c_int(), ctypes.c_ulonglong()
        self._ck(self.lib.wsg_mesh_size(self.h, ctypes.byref(w), ctypes.byref(hh), ctypes.byref(n)))
        return w.value, hh.value, n.value

    def mesh_download(self):
        W, H, _ = self.mesh_size()
        valid = np.empty((H, W), np.uint8)
        p3d = np.empty((H, W, 3), np.float64)
        grey = np.empty((H, W), np.uint8)
        self._ck(self.lib.wsg_mesh_download(self.h, valid.ctypes.data, p3d.ctypes.data, grey.ctypes.data))
        return valid.astype(bool), p3d, grey

    def mesh_zgap_percentile(self, percentile=99.0):
        z = ctypes.c_double()
        self._ck(self.lib.wsg_mesh_zgap_percentile(self.h, percentile, ctypes.byref(z)))
        return z.value

    def mesh_biggest_component(self, zgap):
        n = ctypes.c_ulonglong()
        self._ck(self.lib.wsg_mesh_biggest_component(self.h, zgap, ctypes.byref(n)))
        return n.value

    def mesh_ransac_plane(self, triples, threshold):
        t = np.ascontiguousarray(triples, np.int32)
        plane = (ctypes.c_double * 4)()
        ok, best = ctypes.c_int(), ctypes.c_ulonglong()
        self._ck(self.lib.wsg_mesh_ransac_plane(self.h, t.ctypes.data, t.shape[0], threshold, plane, ctypes.byref(ok), ctypes.byref(best)))
        return bool(ok.value), np.array(list(plane)), best.value

    def mesh_crop_plane(self, plane, threshold):
        n = ctypes.c_ulonglong()
        self._ck(self.lib.wsg_mesh_crop_plane(self.h, _d4(plane), threshold, ctypes.byref(n)))
        return n.value

    def mesh_refine_plane(self, params=None):
        p = params or refine_params()
        plane = (ctypes.c_double * 4)()
        n = ctypes.c_ulonglong()
        self._ck(self.lib.wsg_mesh_refine_plane(self.h, ctypes.byref(p), plane, ctypes.byref(n)))
        return np.array(list(plane)), n.value

    def mesh_refine_inliers(self, params=None, every=10):
        """Every `every`-th refinement inlier in grid scan order (plane_refinement_inliers.xyz): (points[n][3], n_inliers)."""
        p = params or refine_params()
        W, H, _ = self.mesh_size()
        out = np.empty(((W * H + every - 1) // every, 3), np.float64)
        n, nin = ctypes.c_ulonglong(), ctypes.c_ulonglong()
        self._ck(self.lib.wsg_mesh_refine_inliers(self.h, ctypes.byref(p), every, out.ctypes.data, out.shape[0], ctypes.byref(n), ctypes.byref(nin)))
        return out[:n.value].copy(), nin.value

    def mesh_export_xyzc(self, plane, out=None):
        """Bytes of mesh_cam.xyzC.  `out`: optional reusable uint8 destination of at least 148 + 6*W*H bytes (e.g. a numpy
        view of pinned memory); a view of it is returned instead of a fresh bytes object (no allocation, no extra copy)."""
        W, H, _ = self.mesh_size()
        need = 148 + W * H * 6
        buf = out if out is not None else np.empty(need, np.uint8)
        if buf.size < need:
            raise ValueError("destination too small")
        nb = ctypes.c_size_t()
        self._ck(self.lib.wsg_mesh_export_xyzc(self.h, _d4(plane), buf.ctypes.data, buf.size, ctypes.byref(nb)))
        return buf[:nb.value] if out is not None else buf[:nb.value].tobytes()

    def mesh_export_xyzbin(self):
        W, H, _ = self.mesh_size()
        buf = np.empty(4 + W * H * 12, np.uint8)
        nb = ctypes.c_size_t()
        self._ck(self.lib.wsg_mesh_export_xyzbin(self.h, buf.ctypes.data, buf.size, ctypes.byref(nb)))
        return buf[:nb.value].tobytes()

    def xyzc_decode_align(self, xyzc_bytes, align_plane, baseline=1.0):
        """mesh_cam.xyzC bytes -> 3xN float64 points on `align_plane` (load_camera_mesh + align_on_sea_plane * baseline)."""
        buf = np.frombuffer(xyzc_bytes, np.uint8)
        n = int(np.frombuffer(xyzc_bytes[:4], "<u4")[0]) if len(xyzc_bytes) >= 4 else 0
        out = np.empty((3, max(n, 1)), np.float64)
        npts = ctypes.c_size_t()
        self._ck(self.lib.wsg_xyzc_decode_align(self.h, buf.ctypes.data, buf.size, _d4(align_plane), float(baseline),
                                                out.ctypes.data, max(n, 1), ctypes.byref(npts)))
        return out[:, :npts.value].reshape(3, npts.value) if npts.value == max(n, 1) else np.ascontiguousarray(out.reshape(-1)[:3 * npts.value].reshape(3, npts.value))

    def mesh_aligned_points(self, plane, align_plane, baseline=1.0):
        """Device mesh -> 3xN float64 aligned points, bit-identical to xyzc_decode_align(mesh_export_xyzc(plane), ...)."""
        W, H, _ = self.mesh_size()
        out = np.empty(3 * W * H, np.float64)
        npts = ctypes.c_size_t()
        self._ck(self.lib.wsg_mesh_aligned_points(self.h, _d4(plane), _d4(align_plane), float(baseline), out.ctypes.data,
                                                  W * H, ctypes.byref(npts)))
        return out[:3 * npts.value].reshape(3, npts.value).copy()

    def undistort_image(self, img, K, dist):
        """cv::undistort(img, K, dist) as wass_prepare applies it (8-bit grey)."""
        img = np.ascontiguousarray(img, np.uint8)
        H, W = img.shape
        out = np.empty((H, W), np.uint8)
        d = np.ascontiguousarray(np.asarray(dist, np.float64).reshape(-1))
        self._ck(self.lib.wsg_undistort_image(self.h, img.ctypes.data, H, W, W, _darr(K, 9),
                                              d.ctypes.data_as(ctypes.POINTER(ctypes.c_double)) if d.size else None, int(d.size),
                                              out.ctypes.data))
        return out

    def clahe_image(self, img, clip_limit, tiles):
        """cv::createCLAHE(clip_limit, (tiles, tiles)).apply(img) as wass_prepare applies it (8-bit grey)."""
        img = np.ascontiguousarray(img, np.uint8)
        H, W = img.shape
        out = np.empty((H, W), np.uint8)
        self._ck(self.lib.wsg_clahe_image(self.h, img.ctypes.data, H, W, W, float(clip_limit), int(tiles), out.ctypes.data))
        return out

    def prepare_image(self, img, K, dist, clahe_tiles=0, clahe_clip=2.0):
        """process_image() of wass_prepare (wass_prepare.cpp:88-275, no demosaic): CLAHE if clahe_tiles > 0, then undistort."""
        img = np.ascontiguousarray(img, np.uint8)
        H, W = img.shape
        out = np.empty((H, W), np.uint8)
        d = np.ascontiguousarray(np.asarray(dist, np.float64).reshape(-1))
        self._ck(self.lib.wsg_prepare_image(self.h, img.ctypes.data, H, W, W, int(clahe_tiles), float(clahe_clip), _darr(K, 9),
                                            d.ctypes.data_as(ctypes.POINTER(ctypes.c_double)) if d.size else None, int(d.size),
                                            out.ctypes.data))
        return out

    def warp_perspective(self, img, H):
        """cv::warpPerspective(img, H, img.size()) as the custom rectifier applies it (wass_stereo.cpp:515-516)."""
        img = np.ascontiguousarray(img, np.uint8)
        rows, cols = img.shape
        out = np.empty((rows, cols), np.uint8)
        self._ck(self.lib.wsg_warp_perspective(self.h, img.ctypes.data, rows, cols, cols, _darr(H, 9), out.ctypes.data))
        return out

    def rectify_image(self, img, K, Rrect, P):
        img = np.ascontiguousarray(img, np.uint8)
        H, W = img.shape
        out = np.empty((H, W), np.uint8)
        self._ck(self.lib.wsg_rectify_image(self.h, img.ctypes.data, H, W, W, _darr(K, 9), _darr(Rrect, 9), _darr(P, 12), out.ctypes.data))
        return out

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
