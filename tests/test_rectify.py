"""Rectification: wsg_stereo_rectify and wsg_rectify_image against cv2, the library
the reference calls at wass_stereo.cpp:541,600-604."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")


def _rig(rng):
    W, H = int(rng.integers(300, 2600)), int(rng.integers(200, 2100))
    f0, f1 = rng.uniform(0.7, 1.6) * W, rng.uniform(0.7, 1.6) * W
    K0 = np.array([[f0, 0, W / 2 + rng.normal(0, 20)], [0, f0 * rng.uniform(0.98, 1.02), H / 2 + rng.normal(0, 20)], [0, 0, 1]])
    K1 = np.array([[f1, 0, W / 2 + rng.normal(0, 20)], [0, f1 * rng.uniform(0.98, 1.02), H / 2 + rng.normal(0, 20)], [0, 0, 1]])
    R, _ = cv2.Rodrigues(rng.normal(0, 0.08, 3))
    T = np.array([rng.choice([-1, 1]) * 1.0, rng.normal(0, 0.15), rng.normal(0, 0.15)])
    return W, H, K0, K1, R, T / np.linalg.norm(T)


def test_stereo_rectify_matches_cv2():
    from wass_b200 import capi
    rng = np.random.default_rng(0)
    for _ in range(25):
        W, H, K0, K1, R, T = _rig(rng)
        R1, R2, P1, P2, Q, roi1, roi2 = cv2.stereoRectify(K0, np.zeros(5), K1, np.zeros(5), (W, H), R, T.reshape(3, 1),
                                                          flags=0, alpha=1.0, newImageSize=(W, H))
        o = capi.stereo_rectify(K0, K1, R, T, W, H)
        assert np.allclose(o["R1"], R1, atol=1e-12) and np.allclose(o["R2"], R2, atol=1e-12)
        assert np.allclose(o["P1"], P1, rtol=1e-6, atol=1e-6) and np.allclose(o["P2"], P2, rtol=1e-6, atol=1e-6)
        assert o["roi1"] == tuple(roi1) and o["roi2"] == tuple(roi2)


@pytest.mark.gpu
def test_rectify_image_matches_cv2_remap():
    from wass_b200 import capi
    h = capi.Handle(0)
    rng = np.random.default_rng(1)
    for it in range(4):
        W, H, K0, K1, R, T = _rig(rng)
        W, H = min(W, 900), min(H, 700)
        o = capi.stereo_rectify(K0, K1, R, T, W, H)
        img = cv2.resize(rng.integers(0, 256, (H // 5 + 2, W // 5 + 2)).astype(np.float32), (W, H), interpolation=cv2.INTER_CUBIC)
        img = np.clip(img + rng.normal(0, 3, img.shape), 0, 255).astype(np.uint8)
        for K, Rr, P in ((K0, o["R1"], o["P1"]), (K1, o["R2"], o["P2"])):
            m1, m2 = cv2.initUndistortRectifyMap(K, np.zeros(5), Rr, P, (W, H), cv2.CV_32FC1)
            ref = cv2.remap(img, m1, m2, cv2.INTER_CUBIC)
            out = h.rectify_image(img, K, Rr, P)
            diff = np.abs(out.astype(int) - ref.astype(int))
            assert diff.max() <= 1, "max diff %d" % diff.max()
            assert (diff == 0).mean() > 0.999, "only %.4f%% identical" % (100 * (diff == 0).mean())
    h.close()
