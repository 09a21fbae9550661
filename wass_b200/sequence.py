"""In-process sequence runner: the whole per-frame hot path of wass_stereo through the C ABI on ONE handle per GPU
(arena, CUDA context and kernels stay warm across frames), frames round-robin over ranks, one NaN-aware all-reduce of the
plane sums at the end (BASELINE configs[4]).

Per frame, in the order of main() (src/wass_stereo/wass_stereo.cpp:1976-2135):
  sgbm_dense_stereo -> triangulate -> compute_zgap_percentile -> cluster_biggest_connected_component ->
  ransac_find_plane -> crop_plane -> refine_plane -> crop_plane -> save_as_xyz_compressed (kept in memory)

The launcher (wass_b200/launcher.py) does the same through the drop-in executable, one process per frame, as the
reference's drivers do; this module is for callers that already hold rectified images in memory.
"""
import ctypes
import threading
import time
import numpy as np

from . import capi

STAGES = ("dense", "triangulate", "zgap", "component", "ransac", "refine", "export")
_rand_lock = threading.Lock()      # libc's rand() state is process-wide: seed + draw is one critical section


class FrameResult:
    __slots__ = ("plane", "n_points", "xyzc", "ms")

    def __init__(self, plane, n_points, xyzc, ms):
        self.plane, self.n_points, self.xyzc, self.ms = plane, n_points, xyzc, ms


def rectified_calib(K0, K1, R, T, width, height):
    """Calibration dict for wsg_triangulate_from_dense from the rig parameters (cv::stereoRectify on the host)."""
    r = capi.stereo_rectify(K0, K1, R, T, width, height)
    return dict(K0=np.asarray(K0, np.float64), K1=np.asarray(K1, np.float64), R=np.asarray(R, np.float64),
                T=np.asarray(T, np.float64), R1=r["R1"], R2=r["R2"], P1=r["P1"], P2=r["P2"], roi_left=r["roi1"], roi_right=r["roi2"])


def process_frame(h, left_rect, right_rect, calib, dense, ransac_rounds=400, ransac_threshold=1.0, plane_max_distance=1.5,
                  zgap_percentile=99.0, min_points=100, seed=None, keep_xyzc=True, xyzc_out=None):
    """left_rect/right_rect: rectified 8-bit images (full rectified size); calib: dict as rectified_calib returns.
    xyzc_out: optional reusable (pinned) uint8 buffer for the .xyzC bytes; FrameResult.xyzc is then a view of it.
    Returns FrameResult; plane is 4 NaNs when RANSAC fails (the reference writes "nan nan nan nan" and carries on)."""
    ms = {}
    t = time.perf_counter()

    def lap(name):
        nonlocal t
        now = time.perf_counter()
        ms[name] = (now - t) * 1e3
        t = now

    rl, rr = calib["roi_left"], calib["roi_right"]
    lc = np.ascontiguousarray(left_rect[rl[1]:rl[1] + rl[3], rl[0]:rl[0] + rl[2]])
    rc = np.ascontiguousarray(right_rect[rr[1]:rr[1] + rr[3], rr[0]:rr[0] + rr[2]])
    h.dense_stereo(lc, rc, dense, want_host=False)      # the disparity stays on the device for the triangulation
    lap("dense")
    n = h.triangulate_from_dense(left_rect, right_rect, calib, left_rect.shape)
    lap("triangulate")
    if n < min_points:
        raise RuntimeError("too few triangulated points (%d)" % n)
    zg = h.mesh_zgap_percentile(zgap_percentile)
    lap("zgap")
    h.mesh_biggest_component(zg)
    lap("component")
    with _rand_lock:
        if seed is not None:
            ctypes.CDLL("libc.so.6").srand(int(seed))
        draws = capi.ransac_draw(rr[2], rr[3], ransac_rounds)
    ok, plane, _ = h.mesh_ransac_plane(draws, ransac_threshold)
    lap("ransac")
    if ok:
        h.mesh_crop_plane(plane, ransac_threshold)
        plane, _ = h.mesh_refine_plane()
        npts = h.mesh_crop_plane(plane, plane_max_distance)
        lap("refine")
    else:
        plane = np.full(4, np.nan)
        npts = h.mesh_size()[2]
        ms["refine"] = 0.0
    buf = None
    if keep_xyzc:
        # the file needs a plane to rotate into; the reference writes it with the fitted plane
        buf = h.mesh_export_xyzc(plane if ok else np.array([0.0, 0.0, 1.0, 0.0]), out=xyzc_out)
    lap("export")
    return FrameResult(np.asarray(plane, np.float64), int(npts), buf, ms)


def run_sequence(frames, calib, dense, device=0, rank=0, world=1, dist=None, handle=None, xyzc_out=None, **kw):
    """frames: list of callables or (left, right) tuples, one per frame of the WHOLE sequence; this rank processes frames
    rank, rank+world, ...  Returns (mean_plane, planes[n_frames][4], results of the owned frames).

    handle: one capi.Handle or a list of them.  With a list of k handles, k frames are in flight on this GPU, one host
    thread and one stream each (the C calls release the GIL): the aggregation sweeps are latency-bound (DESIGN.md section
    4), so a second frame's kernels fill the SMs' idle issue slots (with three or more handles, cap each one's sweeps at
    half the SMs first: h.sgbm_set_sweep_workers).  Frame i gets seed i either way, so the planes do not
    depend on k.  xyzc_out: a reusable buffer, or a list with one per handle.  Without `handle` one is created and
    destroyed here (the arena is several GB: keep one across calls when processing more than one sequence)."""
    from . import launcher
    own = handle is None
    hs = [capi.Handle(device)] if own else (list(handle) if isinstance(handle, (list, tuple)) else [handle])
    outs = list(xyzc_out) if isinstance(xyzc_out, (list, tuple)) else [xyzc_out] * len(hs)
    if len(hs) > 1 and xyzc_out is not None and len({id(o) for o in outs}) != len(hs):
        raise ValueError("one xyzc_out buffer per handle is needed when frames are in flight concurrently")
    try:
        owned = launcher.shard(len(frames), rank, world)
        results = [None] * len(owned)
        errors = []

        def work(k):
            try:
                for j in range(k, len(owned), len(hs)):
                    i = owned[j]
                    f = frames[i]() if callable(frames[i]) else frames[i]
                    results[j] = process_frame(hs[k], f[0], f[1], calib, dense, seed=i, xyzc_out=outs[k], **kw)
            except BaseException as e:      # re-raised on the caller's thread
                errors.append(e)

        if len(hs) == 1:
            work(0)
        else:
            ths = [threading.Thread(target=work, args=(k,)) for k in range(len(hs))]
            for t in ths:
                t.start()
            for t in ths:
                t.join()
        if errors:
            raise errors[0]
        mean, allp = launcher.reduce_planes([r.plane for r in results], len(frames), owned, dist)
        return mean, allp, results
    finally:
        if own:
            hs[0].close()
