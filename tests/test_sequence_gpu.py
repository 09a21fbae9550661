"""In-process sequence runner (wass_b200/sequence.py): the per-frame order of main() (wass_stereo.cpp:1976-2135) through the
C ABI.  Results must not depend on how many frames are in flight, nor on which handle processed a frame."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _frames(W, H, D, n):
    from wass_b200 import synth
    out = []
    for s in range(n):
        right, left, _ = synth.make_pair(W, H, D, seed=50 + s, d0=10.0 + s)
        out.append((left, right))
    return out


def test_frames_in_flight_do_not_change_results():
    from wass_b200 import capi, sequence, synth
    W, H, D = 640, 400, 64
    c = synth.make_calibration(W, H)
    calib = sequence.rectified_calib(c["K0"], c["K1"], c["R"], c["T"], W, H)
    dense = capi.dense_params(MAX_DISPARITY=D, mode=capi.MODE_SGBM)
    frames = _frames(W, H, D, 5)
    h1 = capi.Handle(0)
    hs = [capi.Handle(0), capi.Handle(0)]
    try:
        m1, p1, r1 = sequence.run_sequence(frames, calib, dense, handle=h1, ransac_rounds=60)
        outs = [[np.empty(148 + 6 * W * H, np.uint8) for _ in range(2)] for _ in hs]
        m2, p2, r2 = sequence.run_sequence(frames, calib, dense, handle=hs, xyzc_out=outs, batch=2, ransac_rounds=60, keep_xyzc=False)
        assert not np.isnan(p1).any()
        np.testing.assert_array_equal(p1, p2)
        np.testing.assert_array_equal(m1, m2)
        assert [r.n_points for r in r1] == [r.n_points for r in r2]
        # the same frame on another handle gives the same bytes
        a = sequence.process_frame(h1, frames[3][0], frames[3][1], calib, dense, seed=3, ransac_rounds=60)
        b = sequence.process_frame(hs[1], frames[3][0], frames[3][1], calib, dense, seed=3, ransac_rounds=60, xyzc_out=outs[1][0])
        assert bytes(a.xyzc) == bytes(b.xyzc) == bytes(r1[3].xyzc)
        np.testing.assert_array_equal(a.plane, p1[3])
        # every plane is a unit normal close to the synthetic fronto-parallel scene's
        assert np.allclose(np.linalg.norm(p1[:, :3], axis=1), 1.0, atol=1e-9)
    finally:
        h1.close()
        for h in hs:
            h.close()


def test_batches_do_not_change_results():
    """process_batch: one batched matcher run for several frames, then the per-frame stages; identical to one frame at a time."""
    from wass_b200 import capi, sequence, synth
    W, H, D = 480, 300, 48
    c = synth.make_calibration(W, H)
    calib = sequence.rectified_calib(c["K0"], c["K1"], c["R"], c["T"], W, H)
    dense = capi.dense_params(MAX_DISPARITY=D, mode=capi.MODE_HH)
    frames = _frames(W, H, D, 5)
    h = capi.Handle(0)
    try:
        m1, p1, r1 = sequence.run_sequence(frames, calib, dense, handle=h, ransac_rounds=50)
        m3, p3, r3 = sequence.run_sequence(frames, calib, dense, handle=h, batch=3, ransac_rounds=50)
        np.testing.assert_array_equal(p1, p3)
        np.testing.assert_array_equal(m1, m3)
        assert [bytes(a.xyzc) for a in r1] == [bytes(b.xyzc) for b in r3]
    finally:
        h.close()


def test_burned_pixels_are_tested_on_the_original_images():
    """DISCARD_BURNED_AREAS and the point colour use the ORIGINAL images at un-rectified coordinates
    (wass_stereo.cpp:1069-1093, 1244-1250, 1342), not the rectified ones the matcher sees."""
    from wass_b200 import capi, sequence, synth
    W, H, D = 480, 300, 48
    c = synth.make_calibration(W, H)
    calib = sequence.rectified_calib(c["K0"], c["K1"], c["R"], c["T"], W, H)
    dense = capi.dense_params(MAX_DISPARITY=D)
    left, right = _frames(W, H, D, 1)[0]
    lo, ro = left.copy(), right.copy()
    ro[100:140, 200:260] = 255                       # saturated patch in the original right image only
    h = capi.Handle(0)
    try:
        a = sequence.process_frame(h, left, right, calib, dense, seed=1, ransac_rounds=40)
        b = sequence.process_frame(h, left, right, calib, dense, seed=1, ransac_rounds=40, left=lo, right=ro)
        assert 40 * 60 * 0.5 < a.n_points - b.n_points <= 40 * 60 + 200     # the patch is gone, nothing else
    finally:
        h.close()


def test_ransac_failure_exports_with_the_best_hypothesis():
    """Soft failure (wass_stereo.cpp:2101-2107, PovMesh.cpp:745-749): plane is NaN, the mesh is still exported with the best
    RANSAC hypothesis -- finite header, decodable points."""
    from wass_b200 import capi, sequence, synth
    from oracle import pipeline as op
    W, H, D = 480, 300, 48
    c = synth.make_calibration(W, H)
    calib = sequence.rectified_calib(c["K0"], c["K1"], c["R"], c["T"], W, H)
    dense = capi.dense_params(MAX_DISPARITY=D)
    left, right = _frames(W, H, D, 1)[0]
    h = capi.Handle(0)
    try:
        r = sequence.process_frame(h, left, right, calib, dense, seed=2, ransac_rounds=30, ransac_threshold=1e-7)
        assert np.isnan(r.plane).all() and r.n_points > 1000
        pts = op.xyz_compressed_decode(bytes(r.xyzc))
        assert pts.shape[0] == r.n_points and np.isfinite(pts).all()
    finally:
        h.close()


def test_shared_output_buffer_is_refused():
    from wass_b200 import capi, sequence, synth
    W, H, D = 320, 200, 32
    c = synth.make_calibration(W, H)
    calib = sequence.rectified_calib(c["K0"], c["K1"], c["R"], c["T"], W, H)
    hs = [capi.Handle(0), capi.Handle(0)]
    try:
        buf = np.empty(148 + 6 * W * H, np.uint8)
        with pytest.raises(ValueError):      # one list of `batch` buffers per handle is needed
            sequence.run_sequence(_frames(W, H, D, 2), calib, capi.dense_params(MAX_DISPARITY=D), handle=hs, xyzc_out=[[buf]])
    finally:
        for h in hs:
            h.close()


def test_sharded_ranks_cover_the_sequence():
    """Two 'ranks' run one after the other on one GPU: the union of their frames is the single-rank result."""
    from wass_b200 import capi, sequence, synth
    W, H, D = 480, 300, 48
    c = synth.make_calibration(W, H)
    calib = sequence.rectified_calib(c["K0"], c["K1"], c["R"], c["T"], W, H)
    dense = capi.dense_params(MAX_DISPARITY=D)
    frames = _frames(W, H, D, 4)
    h = capi.Handle(0)
    try:
        _, pall, _ = sequence.run_sequence(frames, calib, dense, handle=h, ransac_rounds=40, keep_xyzc=False)
        parts = np.full_like(pall, np.nan)
        for r in range(2):
            _, pr, res = sequence.run_sequence(frames, calib, dense, handle=h, rank=r, world=2, ransac_rounds=40, keep_xyzc=False)
            own = list(range(r, 4, 2))
            parts[own] = np.array([x.plane for x in res])
        np.testing.assert_array_equal(parts, pall)
    finally:
        h.close()
