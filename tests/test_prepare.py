"""CLAHE of wass_prepare (src/wass_prepare/wass_prepare.cpp:257-262, 458-462): the oracle's
restatement against cv2.createCLAHE, the kernels against the oracle through the C ABI, and the CLAHE -> undistort chain of
process_image()."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

CASES = [(240, 320, 8, 2.0), (241, 323, 8, 2.0), (300, 400, 150, 2.0), (480, 640, 16, 4.0), (200, 300, 7, 0.0),
         (256, 256, 8, 40.0), (123, 457, 13, 1.5), (64, 64, 64, 2.0), (90, 40, 1, 3.0), (35, 260, 50, 2.0)]


def _img(H, W, seed):
    rng = np.random.default_rng(seed)
    img = cv2.resize(rng.integers(0, 256, (H // 7 + 2, W // 7 + 2), dtype=np.uint8), (W, H), interpolation=cv2.INTER_LINEAR)
    return np.clip(img.astype(int) + rng.integers(-6, 7, (H, W)), 0, 255).astype(np.uint8)


@pytest.mark.parametrize("H,W,tiles,clip", CASES)
def test_oracle_clahe_matches_cv2(H, W, tiles, clip):
    from oracle import pipeline as op
    img = _img(H, W, H + tiles)
    assert np.array_equal(op.clahe_u8(img, clip, tiles), cv2.createCLAHE(clip, (tiles, tiles)).apply(img))


def test_oracle_clahe_flat_and_saturated_images():
    from oracle import pipeline as op
    for img in (np.zeros((64, 96), np.uint8), np.full((64, 96), 255, np.uint8), np.tile(np.arange(96, dtype=np.uint8), (64, 1))):
        assert np.array_equal(op.clahe_u8(img, 2.0, 8), cv2.createCLAHE(2.0, (8, 8)).apply(img))


@pytest.mark.gpu
@pytest.mark.parametrize("H,W,tiles,clip", CASES)
def test_gpu_clahe_matches_oracle(H, W, tiles, clip):
    from wass_b200 import capi
    from oracle import pipeline as op
    img = _img(H, W, H + tiles + 1)
    h = capi.Handle(0)
    try:
        assert np.array_equal(h.clahe_image(img, clip, tiles), op.clahe_u8(img, clip, tiles))
    finally:
        h.close()


@pytest.mark.gpu
def test_gpu_prepare_chain_full_size():
    """process_image(): CLAHE with the documented starting grid of 150, then undistort, at the BASELINE frame size; against
    cv2 on the same host."""
    from wass_b200 import capi
    H, W = 2048, 2448
    img = _img(H, W, 5)
    K = np.array([[2400.0, 0, W / 2 + 3.5], [0, 2410.0, H / 2 - 2.25], [0, 0, 1]])
    dist = np.array([-0.12, 0.05, 1e-3, -5e-4, 0.01])
    h = capi.Handle(0)
    try:
        ref_c = cv2.createCLAHE(2.0, (150, 150)).apply(img)
        got_c = h.clahe_image(img, 2.0, 150)
        assert np.array_equal(got_c, ref_c)
        got = h.prepare_image(img, K, dist, clahe_tiles=150, clahe_clip=2.0)
        assert np.array_equal(got, h.undistort_image(ref_c, K, dist))
        assert np.array_equal(h.prepare_image(img, K, dist), h.undistort_image(img, K, dist))     # CLAHE off: undistort only
    finally:
        h.close()


@pytest.mark.gpu
def test_gpu_clahe_rejects_bad_grid():
    from wass_b200 import capi
    h = capi.Handle(0)
    try:
        with pytest.raises(capi.WsgError):
            h.clahe_image(np.zeros((8, 8), np.uint8), 2.0, 0)
    finally:
        h.close()
