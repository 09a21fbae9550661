"""Pins oracle/pipeline.py (CPU restatement of the stages around the matcher) against cv2 where the
reference calls OpenCV, against hand-computed cases, and against the analytic geometry of the synthetic rig."""
import numpy as np
import pytest
from oracle import pipeline as op


def test_solve3_matches_cv2_solve():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    for _ in range(50):
        M = rng.normal(size=(4, 3))
        A = M.T @ M
        b = rng.normal(size=3)
        ok, x = cv2.solve(A, b.reshape(3, 1), flags=cv2.DECOMP_LU)
        assert ok
        # cv2 4.13 evaluates the 3x3 solve in a slightly different order (max rel. diff seen 1.5e-13)
        assert np.allclose(op.solve3_lu(A, b), x.reshape(3), rtol=1e-11, atol=0)


def test_erode_matches_reference_semantics():
    a = np.ones((6, 7), np.float32)
    a[3, 3] = 0
    e = op.matrix_erode_zero(a)
    exp = np.zeros_like(a)
    exp[1:5, 1:6] = 1
    exp[2:5, 2:5] = 0      # 8-neighbours of the hole
    exp[3, 3] = 0          # centre itself stays what it was (0)
    assert np.array_equal(e, exp)
    # a pixel whose only zero "neighbour" is itself is kept (the centre is not tested, wass_stereo.cpp:690)
    b = np.ones((5, 5), np.float32)
    assert op.matrix_erode_zero(b)[2, 2] == 1


def test_dilate_column_quirk():
    # the ring is centred one column to the right of the written pixel (wass_stereo.cpp:638-659)
    a = np.zeros((5, 8), np.float32)
    a[1, 4] = 6.0
    a[3, 4] = 2.0
    d = op.matrix_dilate_zero(a)
    # written column k uses rows i-1 / i+1 at columns k..k+2 and row i at columns k, k+2
    assert d[2, 2] == 4.0 and d[2, 3] == 4.0 and d[2, 4] == 4.0     # k=2,3,4 see both values
    assert d[2, 5] == 0.0 and d[2, 1] == 0.0
    # needs more than one positive neighbour
    b = np.zeros((5, 8), np.float32)
    b[1, 4] = 6.0
    assert op.matrix_dilate_zero(b).sum() == 6.0
    # last two columns and first/last rows are never written
    c = np.zeros((4, 4), np.float32)
    c[0, :] = 5
    c[2, :] = 5
    dd = op.matrix_dilate_zero(c)
    assert dd[1, 0] == 5 and dd[1, 1] == 5 and dd[1, 2] == 0 and dd[1, 3] == 0


def test_clean_and_convert():
    d16 = np.array([[16, 17, 32, 16 * 64, 16 * 64 + 1, -16, 0]], np.int16)
    out = op.clean_and_convert_disparity(d16, 1, 64, 0, 1.0)
    assert np.array_equal(out, np.array([[0, 17 / 16.0, 2, 64, 0, 0, 0]], np.float32))
    out = op.clean_and_convert_disparity(d16, 1, 64, 3, 1.0)
    assert out[0, 2] == 5.0


def test_rt_from_plane_is_rotation_and_xyzc_roundtrip():
    n = np.array([0.1, -0.5, 0.8])
    n /= np.linalg.norm(n)
    plane = np.array([n[0], n[1], n[2], -3.0])
    R, T, Rinv, Tinv = op.rt_from_plane(*plane)
    assert np.allclose(R @ R.T, np.eye(3), atol=1e-12)
    assert np.allclose(R @ n, [0, 0, 1], atol=1e-12)
    rng = np.random.default_rng(1)
    valid = rng.random((20, 30)) > 0.3
    p3d = rng.normal(size=(20, 30, 3)) * 5 + [0, 0, 20]
    buf = op.xyz_compressed_bytes(valid, p3d, plane)
    assert len(buf) == 148 + 6 * valid.sum()
    dec = op.xyz_compressed_decode(buf)
    span = p3d[valid].max(0) - p3d[valid].min(0)
    assert np.abs(dec - p3d[valid]).max() < 3 * span.max() / 65535.0


def test_identity_rig_matches_cv2_stereoRectify():
    cv2 = pytest.importorskip("cv2")
    from wass_b200 import synth
    W, H = 320, 240
    c = synth.make_calibration(W, H)
    R1, R2, P1, P2, Q, roi1, roi2 = cv2.stereoRectify(c["K0"], np.zeros(5), c["K1"], np.zeros(5), (W, H), c["R"],
                                                      c["T"].reshape(3, 1), flags=0, alpha=1.0, newImageSize=(W, H))
    o = op.rectified_calibration_identity(c["K0"], c["T"], W, H)
    assert np.allclose(R1, o["R1"]) and np.allclose(R2, o["R2"])
    assert np.allclose(P1, o["P1"]) and np.allclose(P2, o["P2"])
    assert P2[0, 3] > 0                        # no auto-swap (wass_stereo.cpp:558-562)
    assert tuple(roi1) == o["roi_left"] and tuple(roi2) == o["roi_right"]


def test_triangulation_depth_is_f_over_d():
    from wass_b200 import synth
    W, H = 64, 48
    c = synth.make_calibration(W, H)
    cal = op.rectified_calibration_identity(c["K0"], c["T"], W, H)
    disp = np.zeros((H, W), np.float32)
    disp[10:40, 20:60] = 8.0
    img = np.full((H, W), 100, np.uint8)
    m = op.triangulate(disp, cal, img, img, min_angle=-1)
    assert m["n"] > 0
    assert m["valid"].shape == (H - 1, W - 1)
    z = m["p3d"][..., 2][m["valid"]]
    assert np.allclose(z, W / 8.0, rtol=1e-9)
    # far points are dropped by the 200-baseline gate, near by z<1 (wass_stereo.cpp:1328-1340)
    disp[:] = 0.2
    assert op.triangulate(disp, cal, img, img, min_angle=-1)["n"] == 0


def test_biggest_component_tie_goes_to_first_in_column_major_order():
    valid = np.zeros((6, 8), bool)
    z = np.zeros((6, 8))
    valid[4:6, 0:2] = True          # 4 px, column-major first
    valid[0:2, 5:7] = True          # 4 px
    valid[0, 0] = True              # 1 px (scanned first, but smaller)
    out = op.biggest_component(valid, z, 1.0)
    assert out[4:6, 0:2].all() and out.sum() == 4
    # z-gap cuts an edge
    valid[:] = False
    valid[2, 0:6] = True
    z[2, 3:] = 10.0
    out = op.biggest_component(valid, z, 1.0)
    assert out.sum() == 3 and out[2, 0:3].all()


def test_zgap_percentile_small():
    valid = np.ones((3, 4), bool)
    z = np.arange(12, dtype=np.float64).reshape(3, 4) ** 2
    g = []
    for i in range(1, 3):
        for j in range(1, 3):
            for dj in (-1, 0, 1):
                g.append(abs(z[i, j] - z[i - 1, j + dj]))
    g = sorted(g)
    assert op.zgap_percentile(valid, z, 50.0) == g[int(np.floor(0.5 * len(g)))]


def test_ransac_and_refine_recover_a_plane():
    rng = np.random.default_rng(3)
    H, W = 40, 60
    v, u = np.mgrid[0:H, 0:W]
    n = np.array([0.05, -0.6, 0.8]); n /= np.linalg.norm(n)
    X = (u - W / 2) * 0.5
    Y = (v - H / 2) * 0.5
    Z = (-(-30.0) - n[0] * X - n[1] * Y) / n[2] + rng.normal(0, 0.01, (H, W))
    p3d = np.stack([X, Y, Z], -1)
    valid = rng.random((H, W)) > 0.1
    rr = op.LibcRand(12345)
    tr = op.ransac_draw_triples(rr, W, H, 50)
    ok, plane, best = op.ransac_find_plane(valid, p3d, tr, 0.1)
    assert ok and best > 0.8 * valid.sum()
    v2 = op.crop_plane(valid, p3d, plane, 0.1)
    plane2, nin = op.refine_plane(v2, p3d)
    assert np.allclose(plane2[:3], n, atol=2e-3) and abs(plane2[3] + 30.0) < 0.05
