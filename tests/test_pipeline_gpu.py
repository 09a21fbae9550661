"""Parity of the CUDA stages around the matcher (through the C ABI) against oracle/pipeline.py.
float32 disparity: bit-exact.  fp64 points: |rel diff| <= 1e-9 (tolerance stated by SURVEY.md §8d; the
arithmetic is ordered like the reference, so most points are in fact identical).  .xyzC: +-1 LSB."""
import ctypes
import os
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
P3D_RTOL = 1e-9


@pytest.fixture(scope="module")
def handle():
    from wass_b200 import capi
    h = capi.Handle(0)
    yield h
    h.close()


def _scene(W, H, D, seed, d0=16.0):
    from wass_b200 import synth
    from oracle import pipeline as op
    right, left, dtrue = synth.make_pair(W, H, D, seed=seed, d0=d0)
    c = synth.make_calibration(W, H)
    cal = op.rectified_calibration_identity(c["K0"], c["T"], W, H)
    return right, left, cal


@pytest.mark.parametrize("W,H,D,mode,off", [(200, 80, 64, 0, 0), (200, 80, 64, 1, 0), (150, 60, 32, 0, 3), (150, 60, 32, 1, -2)])
def test_dense_stage_matches_oracle(handle, W, H, D, mode, off):
    from wass_b200 import capi, synth
    from oracle import sgbm, pipeline as op
    right, left, _ = synth.make_pair(W, H, D, seed=W + mode)
    p = capi.dense_params(MAX_DISPARITY=D, mode=mode, DISPARITY_OFFSET=off)
    out, d16 = handle.dense_stereo(left, right, p, want_disp16=True)
    i1, i2 = synth.pad_for_sgbm(right, left, D, disparity_offset=off)
    ref16 = sgbm.compute(i1, i2, sgbm.wass_params(D, mode=mode))["disp"][:, D:D + W]
    assert np.array_equal(d16, ref16)
    ref = op.postprocess_disparity(ref16, 1, D, disparity_offset=off)
    assert out.dtype == np.float32 and np.array_equal(out, ref)
    assert (out > 0).mean() > 0.2      # small frames lose a lot to the three erosions


@pytest.mark.parametrize("dil,ero", [(0, 0), (1, 2), (3, 1), (2, 0)])
def test_postprocess_steps(handle, dil, ero):
    from oracle import pipeline as op
    rng = np.random.default_rng(dil * 10 + ero)
    d16 = (rng.integers(0, 70 * 16, (57, 91))).astype(np.int16)
    d16[rng.random(d16.shape) < 0.3] = 0
    out = handle.disparity_postprocess(d16, 1, 64, 2, 1.0, dil, ero)
    ref = op.clean_and_convert_disparity(d16, 1, 64, 2, 1.0)
    for _ in range(dil):
        ref = op.matrix_dilate_zero(ref)
    for _ in range(ero):
        ref = op.matrix_erode_zero(ref)
    ref = op.matrix_erode_zero(ref)
    assert np.array_equal(out, ref)


def test_triangulation_matches_oracle(handle):
    from wass_b200 import capi
    from oracle import pipeline as op
    W, H, D = 96, 64, 48
    right, left, cal = _scene(W, H, D, 5)
    p = capi.dense_params(MAX_DISPARITY=D, mode=1)
    roi = cal["roi_right"]
    lc, rc = left[:roi[3], :roi[2]].copy(), right[:roi[3], :roi[2]].copy()
    disp_roi = handle.dense_stereo(lc, rc, p)
    disp = np.zeros((H, W), np.float32)
    disp[:roi[3], :roi[2]] = disp_roi
    ref = op.triangulate(disp, cal, left, right)
    n = handle.triangulate(disp, left, right, cal)
    valid, p3d, grey = handle.mesh_download()
    assert ref["n"] > 500
    assert n == ref["n"] and np.array_equal(valid, ref["valid"])
    assert np.array_equal(grey, ref["color"])
    rel = np.abs(p3d - ref["p3d"])[valid] / np.abs(ref["p3d"][valid]).max()
    assert rel.max() <= P3D_RTOL
    # device-resident hand-off gives the same mesh
    n2 = handle.triangulate_from_dense(left, right, cal, (H, W))
    v2, p2, _ = handle.mesh_download()
    assert n2 == n and np.array_equal(v2, valid) and np.array_equal(p2, p3d)


def test_triangulation_gates(handle):
    from wass_b200 import capi
    from oracle import pipeline as op
    W, H = 80, 50
    right, left, cal = _scene(W, H, 32, 9)
    left = left.copy(); left[10:20, 30:40] = 255      # burned area (wass_stereo.cpp:1069-1074)
    disp = np.zeros((H, W), np.float32)
    disp[5:45, 20:78] = np.linspace(0.5, 30, 58, dtype=np.float32)[None, :]
    lm = np.ones((H, W), np.uint8); lm[30:35, :] = 0
    for kw, okw in [(dict(), dict()), (dict(TRIANG_MIN_ANGLE=-1.0), dict(min_angle=-1.0)),
                    (dict(TRIANG_BBOX_TOP=8.0, TRIANG_BBOX_LEFT=10.0, TRIANG_BBOX_RIGHT=60.0, TRIANG_BBOX_BOTTOM=40.0),
                     dict(bbox=(10.0, 8.0, 60.0, 40.0))), (dict(DISCARD_BURNED_AREAS=0), dict(discard_burned=False))]:
        ref = op.triangulate(disp, cal, left, right, left_mask=lm, **okw)
        n = handle.triangulate(disp, left, right, cal, capi.tri_params(**kw), left_mask=lm)
        valid, p3d, _ = handle.mesh_download()
        assert n == ref["n"] and np.array_equal(valid, ref["valid"])
        if n:
            assert np.abs(p3d - ref["p3d"])[valid].max() <= P3D_RTOL * np.abs(ref["p3d"][valid]).max()


_TRI = np.load(os.path.join(os.path.dirname(__file__), "golden", "triang_golden.npz"))


@pytest.mark.parametrize("name", [str(n) for n in _TRI["names"]])
def test_gpu_triangulation_matches_the_reference_itself(handle, name):
    """The CUDA triangulation against the output of the REFERENCE'S OWN triangulate( StereoMatchEnv& ) on the same inputs
    (tests/golden/triang_golden.npz, made by oracle/_ref/triang_ref: tests/golden/make_triang_golden.py): every gate, both
    un-rectification branches, masks, burned areas, DENSE_SCALE / disparity compensation."""
    from wass_b200 import capi
    g = lambda k: _TRI[name + "/" + k]
    kv = {}
    for line in bytes(g("config")).decode().splitlines():
        if "=" in line:
            k, v = line.split("=", 1)
            kv[k.strip()] = v.strip().strip('"')
    calib = dict(K0=g("K0"), K1=g("K1"), R=g("R"), T=g("T"), R1=g("R1"), R2=g("R2"), P1=g("P1"), P2=g("P2"),
                 roi_left=tuple(int(v) for v in g("roiL")), roi_right=tuple(int(v) for v in g("roiR")))
    if kv.get("USE_CUSTOM_STEREORECTIFY") == "true":
        calib["HLi"], calib["HRi"] = g("HLi"), g("HRi")
    comp, camdist = g("scal")
    tp = capi.tri_params(TRIANG_MIN_ANGLE=float(kv.get("TRIANG_MIN_ANGLE", 20.0)),
                         DISCARD_BURNED_AREAS=int(kv.get("DISCARD_BURNED_AREAS", "true") == "true"),
                         disparity_compensation=int(comp), DENSE_SCALE=float(kv.get("DENSE_SCALE", 1.0)), cam_distance=float(camdist),
                         **{k: float(kv[k]) for k in ("TRIANG_BBOX_TOP", "TRIANG_BBOX_LEFT", "TRIANG_BBOX_RIGHT", "TRIANG_BBOX_BOTTOM") if k in kv})
    n = handle.triangulate(g("disp"), g("left"), g("right"), calib, tp,
                           left_mask=g("lmask") if "LEFT_MASK_IMAGE" in kv else None,
                           right_mask=g("rmask") if "RIGHT_MASK_IMAGE" in kv else None)
    valid, p3d, grey = handle.mesh_download()
    ref_valid = g("valid").astype(bool)
    assert n == int(g("n")[0]) and np.array_equal(valid.astype(bool), ref_valid)
    assert np.abs(p3d - g("xyz"))[ref_valid].max() <= P3D_RTOL * np.abs(g("xyz")[ref_valid]).max()
    assert np.array_equal(grey[ref_valid], g("grey")[ref_valid])


def _plane_mesh(H, W, seed, holes=0.1):
    rng = np.random.default_rng(seed)
    v, u = np.mgrid[0:H, 0:W]
    n = np.array([0.05, -0.6, 0.8]); n /= np.linalg.norm(n)
    X = (u - W / 2) * 0.11
    Y = (v - H / 2) * 0.09
    Z = (30.0 - n[0] * X - n[1] * Y) / n[2] + rng.normal(0, 0.02, (H, W))
    Z[H // 3:H // 3 + 4, :] += 3.0          # a ledge: z-gap outliers
    p3d = np.stack([X, Y, Z], -1)
    valid = rng.random((H, W)) > holes
    valid[:, W // 2] = False                # splits the grid into two components
    valid[H // 2, W // 2] = True
    return valid, p3d, rng.integers(0, 255, (H, W)).astype(np.uint8)


def test_mesh_ops_match_oracle(handle):
    from wass_b200 import capi
    from oracle import pipeline as op
    H, W = 70, 101
    valid, p3d, grey = _plane_mesh(H, W, 4)
    handle.mesh_upload(valid, p3d, grey)
    assert handle.mesh_size() == (W, H, int(valid.sum()))
    v0, p0, g0 = handle.mesh_download()
    assert np.array_equal(v0, valid) and np.array_equal(p0, p3d) and np.array_equal(g0, grey)
    for pct in (50.0, 90.0, 99.0):
        assert handle.mesh_zgap_percentile(pct) == op.zgap_percentile(valid, p3d[..., 2], pct)
    zg = op.zgap_percentile(valid, p3d[..., 2], 99.0)
    ref_v = op.biggest_component(valid, p3d[..., 2], zg)
    nleft = handle.mesh_biggest_component(zg)
    v1, _, _ = handle.mesh_download()
    assert np.array_equal(v1, ref_v) and nleft == ref_v.sum()
    # RANSAC with the same host-drawn triples
    ctypes.CDLL("libc.so.6").srand(4242)
    tr = capi.ransac_draw(W, H, 60)
    ok, plane, best = handle.mesh_ransac_plane(tr, 0.05)
    rok, rplane, rbest = op.ransac_find_plane(ref_v, p3d, tr, 0.05)
    assert ok == rok and best == rbest
    assert np.allclose(plane, rplane, rtol=1e-9, atol=1e-12)
    nl = handle.mesh_crop_plane(plane, 0.05)
    ref_v2 = op.crop_plane(ref_v, p3d, rplane, 0.05)
    v2, _, _ = handle.mesh_download()
    assert nl == ref_v2.sum() and np.array_equal(v2, ref_v2)
    # the sample main() dumps to plane_refinement_inliers.xyz, selected on the device
    for every, ct in ((10, False), (1, False), (7, True)):
        rp = capi.refine_params(PLANE_USE_CENTRAL_THIRD_ONLY=int(ct))
        smp, nin0 = handle.mesh_refine_inliers(rp, every)
        rsmp, rnin0 = op.refinement_inlier_samples(ref_v2, p3d, every, central_third=ct)
        assert nin0 == rnin0 and smp.shape == rsmp.shape and np.array_equal(smp, rsmp)
    plane2, nin = handle.mesh_refine_plane()
    rplane2, rnin = op.refine_plane(ref_v2, p3d)
    assert nin == rnin and np.allclose(plane2, rplane2, rtol=1e-6, atol=1e-9)
    handle.mesh_crop_plane(plane2, 1.5)
    ref_v3 = op.crop_plane(ref_v2, p3d, rplane2, 1.5)
    # exports
    buf = handle.mesh_export_xyzc(plane2)
    ref = op.xyz_compressed_bytes(ref_v3, p3d, plane2)
    assert len(buf) == len(ref) and buf[:4] == ref[:4]
    hdr_a, hdr_b = np.frombuffer(buf[4:148], "<f8"), np.frombuffer(ref[4:148], "<f8")
    assert np.allclose(hdr_a, hdr_b, rtol=1e-12, atol=1e-15)
    qa, qb = np.frombuffer(buf[148:], "<u2").astype(int), np.frombuffer(ref[148:], "<u2").astype(int)
    assert np.abs(qa - qb).max() <= 1
    dec = op.xyz_compressed_decode(buf)            # the consumer-side reader (wass_utils.py:22-35)
    span = np.ptp(p3d[ref_v3], axis=0).max()
    assert np.abs(dec - p3d[ref_v3]).max() < 3 * span / 65535.0
    xb = handle.mesh_export_xyzbin()
    assert np.frombuffer(xb[:4], "<u4")[0] == ref_v3.sum()
    assert np.array_equal(np.frombuffer(xb[4:], "<f4").reshape(-1, 3), p3d[ref_v3].astype(np.float32))


def test_biggest_component_tie_rule(handle):
    from oracle import pipeline as op
    valid = np.zeros((9, 12), bool)
    valid[6:8, 0:2] = True
    valid[0:2, 7:9] = True
    valid[0, 0] = True
    p3d = np.zeros((9, 12, 3))
    handle.mesh_upload(valid, p3d)
    assert handle.mesh_biggest_component(1.0) == 4
    v, _, _ = handle.mesh_download()
    assert np.array_equal(v, op.biggest_component(valid, p3d[..., 2], 1.0)) and v[6:8, 0:2].all()


def test_ransac_failure_is_soft(handle):
    valid = np.zeros((30, 40), bool)
    valid[0:3, 0:3] = True
    handle.mesh_upload(valid, np.random.default_rng(0).normal(size=(30, 40, 3)))
    tr = np.array([[0, 0, 2, 2, 1, 0]] * 4, np.int32)
    ok, plane, best = handle.mesh_ransac_plane(tr, 0.5)
    assert not ok and best < 30 * 40 // 10


def test_full_frame_end_to_end_properties(handle):
    """BASELINE-size frame (2448x2048, D=256, 8 paths): size-independent properties instead of an oracle run."""
    from wass_b200 import capi, synth
    from oracle import pipeline as op
    W, H, D = 2448, 2048, 256
    right, left, dtrue = synth.make_pair(W, H, D, seed=0, d0=16.0)
    c = synth.make_calibration(W, H)
    cal = op.rectified_calibration_identity(c["K0"], c["T"], W, H)
    roi = cal["roi_right"]
    p = capi.dense_params(MAX_DISPARITY=D, mode=1)
    disp = handle.dense_stereo(left[:roi[3], :roi[2]].copy(), right[:roi[3], :roi[2]].copy(), p)
    st = handle.sgbm_stats()
    assert st["out_of_domain"] == 0
    ok = disp > 0
    assert ok.mean() > 0.85
    err = np.abs(disp - dtrue[:roi[3], :roi[2]])[ok]
    assert np.median(err) < 0.25 and (err < 1.0).mean() > 0.97          # against the generator's ground truth
    n = handle.triangulate_from_dense(left, right, cal, (H, W))
    assert n > 3_000_000                                                # test/verify_meshes.m:8
    zg = handle.mesh_zgap_percentile(99.0)
    n1 = handle.mesh_biggest_component(zg)
    assert 0.9 * n < n1 <= n
    assert handle.mesh_biggest_component(zg) == n1                      # idempotent
    ctypes.CDLL("libc.so.6").srand(1)
    okp, plane, best = handle.mesh_ransac_plane(capi.ransac_draw(roi[2], roi[3], 100), 1.0)
    assert okp and abs(np.linalg.norm(plane[:3]) - 1) < 1e-12
    handle.mesh_crop_plane(plane, 1.0)
    plane2, nin = handle.mesh_refine_plane()
    n2 = handle.mesh_crop_plane(plane2, 1.5)
    buf = handle.mesh_export_xyzc(plane2)
    assert len(buf) == 148 + 6 * n2
    valid, p3d, _ = handle.mesh_download()
    dec = op.xyz_compressed_decode(buf)
    assert np.abs(dec - p3d[valid]).max() < 3 * np.ptp(p3d[valid], axis=0).max() / 65535.0
    z = p3d[..., 2][valid]
    ztrue = (W / dtrue[:roi[3], :roi[2]])[valid]
    assert np.median(np.abs(z - ztrue) / ztrue) < 0.01
