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


def _mesh_stages(h, roi_right, ransac_rounds, ransac_threshold, plane_max_distance, zgap_percentile, seed, keep_xyzc, xyzc_out,
                 ms, lap):
    """compute_zgap_percentile .. save_as_xyz_compressed on the mesh the handle holds (wass_stereo.cpp:2046-2135)."""
    zg = h.mesh_zgap_percentile(zgap_percentile)
    lap("zgap")
    h.mesh_biggest_component(zg)
    lap("component")
    with _rand_lock:
        if seed is not None:
            ctypes.CDLL("libc.so.6").srand(int(seed))
        draws = capi.ransac_draw(roi_right[2], roi_right[3], ransac_rounds)
    ok, ransac_plane, _ = h.mesh_ransac_plane(draws, ransac_threshold)
    lap("ransac")
    if ok:
        h.mesh_crop_plane(ransac_plane, ransac_threshold)
        plane, _ = h.mesh_refine_plane()
        npts = h.mesh_crop_plane(plane, plane_max_distance)
        export_plane = plane
        lap("refine")
    else:
        # soft failure (wass_stereo.cpp:2101-2107): plane.txt says nan, but the mesh is still written -- with the best
        # RANSAC hypothesis, which is what PovMesh keeps in plane_coeffs when ransac_find_plane returns false
        # (PovMesh.cpp:745-749)
        plane = np.full(4, np.nan)
        export_plane = ransac_plane
        npts = h.mesh_size()[2]
        ms["refine"] = 0.0
    buf = h.mesh_export_xyzc(export_plane, out=xyzc_out) if keep_xyzc else None
    lap("export")
    return np.asarray(plane, np.float64), int(npts), buf


def _clock():
    ms = {}
    t = [time.perf_counter()]

    def lap(name):
        now = time.perf_counter()
        ms[name] = ms.get(name, 0.0) + (now - t[0]) * 1e3
        t[0] = now
    return ms, lap


def process_frame(h, left_rect, right_rect, calib, dense, ransac_rounds=400, ransac_threshold=1.0, plane_max_distance=1.5,
                  zgap_percentile=99.0, min_points=100, seed=None, keep_xyzc=True, xyzc_out=None, left=None, right=None):
    """left_rect/right_rect: rectified 8-bit images (full rectified size), used for the ROI crops the matcher sees.
    left/right: the ORIGINAL (undistorted, un-rectified) images: triangulate() tests DISCARD_BURNED_AREAS and takes the
    point colour at un-rectified coordinates on env.left / env.right (wass_stereo.cpp:1069-1093, 1244-1250, 1342).  They
    default to the rectified pair, which is only right for a rig whose rectification is the identity (the synthetic rig).
    calib: dict as rectified_calib returns.  xyzc_out: optional reusable (pinned) uint8 buffer for the .xyzC bytes.
    Returns FrameResult; plane is 4 NaNs when RANSAC fails (the reference writes "nan nan nan nan" and carries on)."""
    return process_batch(h, [(left_rect, right_rect, left, right)], calib, dense, ransac_rounds, ransac_threshold, plane_max_distance,
                         zgap_percentile, min_points, [seed], keep_xyzc, [xyzc_out])[0]


def process_batch(h, frames, calib, dense, ransac_rounds=400, ransac_threshold=1.0, plane_max_distance=1.5, zgap_percentile=99.0,
                  min_points=100, seeds=None, keep_xyzc=True, xyzc_out=None):
    """frames: list of (left_rect, right_rect[, left, right]) of one size.  ONE batched dense-matcher run for all of them
    (the aggregation sweeps walk the bands of all frames in one launch each), then the per-frame stages.
    xyzc_out: None or one reusable buffer per frame (the returned FrameResult.xyzc are views of them)."""
    n = len(frames)
    seeds = seeds if seeds is not None else [None] * n
    outs = xyzc_out if xyzc_out is not None else [None] * n
    rl, rr = calib["roi_left"], calib["roi_right"]
    ms0, lap0 = _clock()
    lcs = [np.ascontiguousarray(f[0][rl[1]:rl[1] + rl[3], rl[0]:rl[0] + rl[2]]) for f in frames]
    rcs = [np.ascontiguousarray(f[1][rr[1]:rr[1] + rr[3], rr[0]:rr[0] + rr[2]]) for f in frames]
    h.dense_stereo_batch(lcs, rcs, dense)                # the disparities stay on the device for the triangulation
    lap0("dense")
    results = []
    for i, f in enumerate(frames):
        ms, lap = _clock()
        ms["dense"] = ms0["dense"] / n
        left = f[2] if len(f) > 2 and f[2] is not None else f[0]
        right = f[3] if len(f) > 3 and f[3] is not None else f[1]
        h.dense_select(i)
        npts = h.triangulate_from_dense(left, right, calib, f[0].shape)
        lap("triangulate")
        if npts < min_points:
            raise RuntimeError("too few triangulated points (%d)" % npts)
        plane, npts, buf = _mesh_stages(h, rr, ransac_rounds, ransac_threshold, plane_max_distance, zgap_percentile, seeds[i],
                                        keep_xyzc, outs[i], ms, lap)
        results.append(FrameResult(plane, npts, buf, ms))
    return results


def run_sequence(frames, calib, dense, device=0, rank=0, world=1, dist=None, handle=None, xyzc_out=None, batch=1, **kw):
    """frames: list of callables or (left_rect, right_rect[, left, right]) tuples, one per frame of the WHOLE sequence;
    this rank processes frames rank, rank+world, ...  Returns (mean_plane, planes[n_frames][4], results of the owned frames).

    handle: one capi.Handle or a list of them; every handle works through its share of the owned frames `batch` at a time
    (one batched matcher run per `batch` frames: process_batch), one host thread and one stream per handle (the C calls
    release the GIL), so the copies and mesh stages of one batch overlap the matcher of another.  Frame i gets seed i
    either way, so the planes depend on neither.  xyzc_out: None, or a list with `batch` reusable buffers per handle.
    Without `handle` one is created and destroyed here (the arena is several GB per frame of a batch: keep one across
    calls when processing more than one sequence).
    dist: a torch.distributed module for the final plane reduction, or None; see also capi.Handle.plane_allreduce for the
    same reduction through the C ABI's own NCCL communicator."""
    from . import launcher
    own = handle is None
    hs = [capi.Handle(device)] if own else (list(handle) if isinstance(handle, (list, tuple)) else [handle])
    batch = max(1, int(batch))
    if xyzc_out is None:
        outs = [None] * len(hs)
    else:
        outs = list(xyzc_out)
        if len(outs) != len(hs) or any(len(o) < batch for o in outs):
            raise ValueError("xyzc_out needs one list of `batch` buffers per handle")
    try:
        owned = launcher.shard(len(frames), rank, world)
        results = [None] * len(owned)
        errors = []
        chunks = [list(range(j, min(j + batch, len(owned)))) for j in range(0, len(owned), batch)]

        def work(k):
            try:
                for c in chunks[k::len(hs)]:
                    fr = []
                    for j in c:
                        f = frames[owned[j]]
                        fr.append(f() if callable(f) else f)
                    res = process_batch(hs[k], fr, calib, dense, seeds=[owned[j] for j in c],
                                        xyzc_out=None if outs[k] is None else outs[k][:len(c)], **kw)
                    for j, r in zip(c, res):
                        results[j] = r
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
