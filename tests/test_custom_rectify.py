"""USE_CUSTOM_STEREORECTIFY (src/wass_stereo/stereorectify.cpp:57-244, wass_stereo.cpp:299-305, 496-529; SURVEY section 8f
rank 3).  cv2 does not export cv::DownhillSolver, so the reference's optimiser cannot be run here: the optimum is checked
against an independent minimiser and by its properties, everything else (homographies for a given angle, ROI,
cv::warpPerspective, the homography unrectification inside the triangulation) against the oracle / cv2."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")


def _rig(seed, W=640, H=480):
    rng = np.random.default_rng(seed)
    f = W * (0.9 + 0.3 * rng.random())
    K0 = np.array([[f, 0, W / 2 + rng.normal(0, 5)], [0, f * 1.01, H / 2 + rng.normal(0, 5)], [0, 0, 1]])
    K1 = np.array([[f * 1.02, 0, W / 2 + rng.normal(0, 5)], [0, f * 1.03, H / 2 + rng.normal(0, 5)], [0, 0, 1]])
    rv = rng.normal(0, 0.03, 3)
    R = cv2.Rodrigues(rv)[0]
    T = np.array([1.0, rng.normal(0, 0.05), rng.normal(0, 0.05)])
    return K0, K1, R, T, W, H


@pytest.mark.parametrize("seed", range(6))
def test_given_angle_matches_oracle(seed):
    from wass_b200 import capi
    from oracle import pipeline as op
    K0, K1, R, T, W, H = _rig(seed)
    for ang in (3.5, -7.25, 15.0):
        got = capi.stereo_rectify_custom(K0, K1, R, T, W, H, rot_angle=ang)
        H0, H1, roi = op.stereo_rectify_custom(K0, K1, R, T, W, H, ang)
        assert got["angle"] == ang
        np.testing.assert_allclose(got["H0"], H0, rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(got["H1"], H1, rtol=1e-9, atol=1e-12)
        assert got["roi"] == roi
        assert abs(np.linalg.det(got["H0"]) - 1) < 1e-9 and abs(np.linalg.det(got["H1"]) - 1) < 1e-9


@pytest.mark.parametrize("seed", range(6))
def test_optimised_angle_is_the_minimum(seed):
    from wass_b200 import capi
    from oracle import pipeline as op
    K0, K1, R, T, W, H = _rig(10 + seed)
    got = capi.stereo_rectify_custom(K0, K1, R, T, W, H)            # rot_angle == 0: optimise
    f = op._custom_rect_functional(K0, K1, R, T)
    ref = op.custom_rectify_best_angle(K0, K1, R, T)
    assert abs(got["angle"] - ref) < 1e-5, (got["angle"], ref)
    v = f(got["angle"])[0]
    assert v <= f(got["angle"] + 1e-3)[0] and v <= f(got["angle"] - 1e-3)[0] and v <= f(0.0)[0]
    H0, H1, roi = op.stereo_rectify_custom(K0, K1, R, T, W, H, got["angle"])
    np.testing.assert_allclose(got["H0"], H0, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(got["H1"], H1, rtol=1e-9, atol=1e-12)
    assert got["roi"] == roi
    # rectification property: corresponding points end up on the same row
    rng = np.random.default_rng(seed)
    X1 = np.stack([rng.uniform(-3, 3, 50), rng.uniform(-2, 2, 50), rng.uniform(8, 40, 50)])     # camera-1 frame
    X0 = R @ X1 + T[:, None]                                                                       # camera-0 frame
    p0, p1 = K0 @ X0, K1 @ X1
    q0, q1 = got["H0"] @ (p0 / p0[2]), got["H1"] @ (p1 / p1[2])
    assert np.abs(q0[1] / q0[2] - q1[1] / q1[2]).max() < 1e-6


def test_oracle_warp_matches_cv2():
    from oracle import pipeline as op
    rng = np.random.default_rng(3)
    for t in range(6):
        Hh, W = int(rng.integers(60, 400)), int(rng.integers(100, 600))
        img = cv2.resize(rng.integers(0, 256, (Hh // 6 + 2, W // 6 + 2), dtype=np.uint8), (W, Hh), interpolation=cv2.INTER_CUBIC)
        Hm = np.eye(3) + rng.normal(0, 1, (3, 3)) * np.array([[0.05, 0.05, 8], [0.05, 0.05, 8], [1e-4, 1e-4, 0]])
        assert np.array_equal(op.warp_perspective_u8(img, Hm), cv2.warpPerspective(img, Hm, (W, Hh)))


@pytest.mark.gpu
def test_gpu_warp_matches_cv2():
    from wass_b200 import capi, synth
    from oracle import pipeline as op
    rng = np.random.default_rng(4)
    h = capi.Handle(0)
    try:
        for t in range(6):
            Hh, W = int(rng.integers(60, 400)), int(rng.integers(100, 600))
            img = cv2.resize(rng.integers(0, 256, (Hh // 6 + 2, W // 6 + 2), dtype=np.uint8), (W, Hh), interpolation=cv2.INTER_CUBIC)
            Hm = np.eye(3) + rng.normal(0, 1, (3, 3)) * np.array([[0.05, 0.05, 8], [0.05, 0.05, 8], [1e-4, 1e-4, 0]])
            assert np.array_equal(h.warp_perspective(img, Hm), op.warp_perspective_u8(img, Hm))
        # the benchmark frame size with the homographies of a real rig
        right, left, _ = synth.make_pair(2448, 2048, 256, seed=1, d0=16.0)
        K0, K1, R, T, W, Hh = _rig(2, 2448, 2048)
        r = capi.stereo_rectify_custom(K0, K1, R, T, W, Hh)
        for img, Hm in ((left, r["H0"]), (right, r["H1"])):
            assert np.array_equal(h.warp_perspective(img, Hm), cv2.warpPerspective(img, Hm, (W, Hh)))
    finally:
        h.close()


@pytest.mark.gpu
def test_gpu_triangulation_through_homographies():
    """Identity-like check of the homography unrectification: with H = P[:, :3] R_rect K^-1 the custom mode must give the
    points of the standard mode (wass_stereo.cpp:299-322 are two forms of the same map then)."""
    from wass_b200 import capi, synth
    from oracle import pipeline as op
    W, H, D = 320, 200, 48
    right, left, _ = synth.make_pair(W, H, D, seed=9, d0=8.0)
    c = synth.make_calibration(W, H)
    cal = op.rectified_calibration_identity(c["K0"], c["T"], W, H)
    h = capi.Handle(0)
    try:
        roi = cal["roi_right"]
        h.dense_stereo(left[:roi[3], :roi[2]].copy(), right[:roi[3], :roi[2]].copy(), capi.dense_params(MAX_DISPARITY=D))
        n_std = h.triangulate_from_dense(left, right, cal, (H, W))
        v0, p0, _ = h.mesh_download()
        cal2 = dict(cal)
        HL = cal["P1"][:, :3] @ cal["R1"] @ np.linalg.inv(c["K0"])
        HR = cal["P2"][:, :3] @ cal["R2"] @ np.linalg.inv(c["K1"])
        cal2["HLi"], cal2["HRi"] = np.linalg.inv(HL), np.linalg.inv(HR)
        n_h = h.triangulate_from_dense(left, right, cal2, (H, W))
        v1, p1, _ = h.mesh_download()
        assert n_std == n_h and np.array_equal(v0, v1)
        assert np.abs(p0[v0] - p1[v1]).max() < 1e-8 * np.abs(p0[v0]).max()
    finally:
        h.close()
