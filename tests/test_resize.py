"""cv::resize as the DENSE_SCALE != 1 path of sgbm_dense_stereo uses it (wass_stereo.cpp:788-797, 903-904): the oracle's
restatement against cv2 with Intel IPP off (OpenCV's own code, which is what the reference's conda-forge libopencv runs),
and the CUDA kernels against the oracle through the C ABI."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")


@pytest.fixture
def no_ipp():
    opt = cv2.useOptimized()
    cv2.setUseOptimized(False)
    cv2.ipp.setUseIPP(False)
    yield
    cv2.setUseOptimized(opt)


SCALES = [0.25, 0.33, 0.5, 0.7, 0.9, 1.25, 1.5, 2.0, 3.0]


def _u8(rng, h, w, smooth):
    if smooth:
        return cv2.resize(rng.integers(0, 256, (h // 5 + 2, w // 5 + 2), dtype=np.uint8), (w, h), interpolation=cv2.INTER_LINEAR)
    return rng.integers(0, 256, (h, w), dtype=np.uint8)


def _disp(rng, h, w):
    d = (rng.random((h, w)) * 200).astype(np.float32)
    d[rng.random((h, w)) < 0.2] = 0
    return d


def test_oracle_u8_cubic_matches_cv2(no_ipp):
    from oracle import pipeline as op
    rng = np.random.default_rng(11)
    for t in range(24):
        h, w = int(rng.integers(20, 260)), int(rng.integers(20, 420))
        img = _u8(rng, h, w, t % 2)
        s = SCALES[t % len(SCALES)]
        fy = s if s < 1 else 1.0
        ref = cv2.resize(img, None, fx=s, fy=fy, interpolation=cv2.INTER_CUBIC)
        out = op.resize_cubic_u8(img, s, fy)
        assert out.shape == ref.shape and np.array_equal(out, ref), (h, w, s)
        assert np.array_equal(op.dense_input_resize(img, s), ref)


def test_oracle_f32_resizes_match_cv2(no_ipp):
    from oracle import pipeline as op
    rng = np.random.default_rng(12)
    for t in range(24):
        h, w = int(rng.integers(20, 260)), int(rng.integers(20, 420))
        d = _disp(rng, h, w)
        s = SCALES[t % len(SCALES)]
        dw = int(round(w / s)) + int(rng.integers(-2, 3))
        dh = (int(round(h / s)) + int(rng.integers(-2, 3))) if s < 1 else h
        ref = cv2.resize(d, (dw, dh), interpolation=cv2.INTER_CUBIC)
        out = op.resize_cubic_f32(d, dw, dh)
        assert np.array_equal(out.view(np.uint32), ref.view(np.uint32)), (h, w, s)
        assert np.array_equal(op.resize_nearest(d, dw, dh), cv2.resize(d, (dw, dh), interpolation=cv2.INTER_NEAREST))


def test_ipp_path_of_the_wheel_differs(no_ipp):
    """Why parity is defined with IPP off: the wheel's default INTER_CUBIC is Intel IPP's and is not OpenCV's arithmetic."""
    from oracle import pipeline as op
    rng = np.random.default_rng(13)
    img = _u8(rng, 200, 300, True)
    own = op.resize_cubic_u8(img, 0.7, 0.7)
    cv2.setUseOptimized(True)
    if not cv2.ipp.useIPP():
        pytest.skip("this cv2 build has no IPP")
    ipp = cv2.resize(img, None, fx=0.7, fy=0.7, interpolation=cv2.INTER_CUBIC)
    diff = np.abs(own.astype(int) - ipp.astype(int))
    assert diff.max() <= 1 and 0 < (diff > 0).mean() < 0.15


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(6))
def test_gpu_resizes_match_oracle(seed):
    from wass_b200 import capi
    from oracle import pipeline as op
    rng = np.random.default_rng(100 + seed)
    h = capi.Handle(0)
    try:
        for t in range(6):
            H, W = int(rng.integers(20, 300)), int(rng.integers(20, 500))
            s = SCALES[(seed + t) % len(SCALES)]
            fy = s if s < 1 else 1.0
            img = _u8(rng, H, W, t % 2)
            assert np.array_equal(h.resize_u8_cubic(img, s, fy), op.resize_cubic_u8(img, s, fy)), (H, W, s)
            d = _disp(rng, H, W)
            dw = int(round(W / s)) + int(rng.integers(-2, 3))
            dh = (int(round(H / s)) + int(rng.integers(-2, 3))) if s < 1 else H
            assert np.array_equal(h.resize_f32(d, dh, dw, "cubic").view(np.uint32), op.resize_cubic_f32(d, dw, dh).view(np.uint32))
            assert np.array_equal(h.resize_f32(d, dh, dw, "nearest"), op.resize_nearest(d, dw, dh))
    finally:
        h.close()


@pytest.mark.gpu
def test_gpu_resize_full_size():
    """BASELINE frame size, through cv2 (IPP off) when it is importable on the box, else the oracle."""
    from wass_b200 import capi, synth
    from oracle import pipeline as op
    right, _, _ = synth.make_pair(2448, 2048, 256, seed=3, d0=16.0)
    h = capi.Handle(0)
    try:
        for s in (0.5, 0.75, 1.5):
            fy = s if s < 1 else 1.0
            assert np.array_equal(h.resize_u8_cubic(right, s, fy), op.resize_cubic_u8(right, s, fy))
    finally:
        h.close()


@pytest.mark.gpu
@pytest.mark.parametrize("scale,mode", [(0.5, 0), (0.75, 1), (1.5, 0), (2.0, 1)])
def test_dense_stage_with_dense_scale(scale, mode):
    """sgbm_dense_stereo at DENSE_SCALE != 1 (wass_stereo.cpp:764-1020) end to end against the oracle."""
    from wass_b200 import capi, synth
    from oracle import sgbm, pipeline as op
    W, H, D = 240, 100, 64
    right, left, _ = synth.make_pair(W, H, D, seed=7 + mode, d0=6.0)
    h = capi.Handle(0)
    try:
        p = capi.dense_params(MAX_DISPARITY=D, mode=mode, DENSE_SCALE=scale)
        out, d16 = h.dense_stereo(left, right, p, want_disp16=True)
        ls, rs = op.dense_input_resize(left, scale), op.dense_input_resize(right, scale)
        i1, i2 = synth.pad_for_sgbm(rs, ls, D)
        ref16 = sgbm.compute(i1, i2, sgbm.wass_params(D, mode=mode))["disp"][:, D:D + rs.shape[1]]
        assert d16.shape == ref16.shape and np.array_equal(d16, ref16)
        ref = op.postprocess_disparity(ref16, 1, D, dense_scale=scale, out_size=(H, W))
        assert out.shape == (H, W) and np.array_equal(out.view(np.uint32), ref.view(np.uint32))
        # the map is in full-resolution pixels again: close to the generator's field where valid
        assert (out > 0).mean() > 0.1
        # the stand-alone post-filter entry point gives the same result from the int16 map
        out2 = h.disparity_postprocess(ref16, 1, D, dense_scale=scale, out_size=(H, W))
        assert np.array_equal(out2.view(np.uint32), ref.view(np.uint32))
    finally:
        h.close()


def test_dense_scaled_size_matches_cv2_shapes(no_ipp):
    """wsg_dense_scaled_size (host only): the matcher-input size of wass_stereo.cpp:788-797 is cv::resize's own dsize."""
    import ctypes
    from wass_b200 import capi
    lib = capi.load()
    rng = np.random.default_rng(5)
    for _ in range(40):
        H, W = int(rng.integers(3, 400)), int(rng.integers(3, 600))
        s = float(rng.choice([0.25, 0.3, 0.5, 0.55, 0.75, 0.9, 1.0, 1.1, 1.5, 2.0, 2.5]))
        hs, ws = ctypes.c_int(), ctypes.c_int()
        lib.wsg_dense_scaled_size(H, W, s, ctypes.byref(hs), ctypes.byref(ws))
        if s == 1.0:
            assert (hs.value, ws.value) == (H, W)
            continue
        ref = cv2.resize(np.zeros((H, W), np.uint8), None, fx=s, fy=s if s < 1 else 1.0, interpolation=cv2.INTER_CUBIC)
        assert (hs.value, ws.value) == ref.shape, (H, W, s)


def test_oracle_resizes_extreme_scales_and_thin_images(no_ipp):
    from oracle import pipeline as op
    rng = np.random.default_rng(21)
    for (H, W, s) in [(40, 300, 0.1), (300, 40, 0.15), (9, 9, 4.0), (1, 50, 2.0), (50, 1, 0.5), (2, 2, 3.0), (17, 33, 0.07)]:
        img = rng.integers(0, 256, (H, W), dtype=np.uint8)
        fy = s if s < 1 else 1.0
        if round(H * fy) < 1 or round(W * s) < 1:
            continue
        ref = cv2.resize(img, None, fx=s, fy=fy, interpolation=cv2.INTER_CUBIC)
        assert np.array_equal(op.resize_cubic_u8(img, s, fy), ref), (H, W, s)
        d = (rng.random((H, W)) * 100).astype(np.float32)
        dh, dw = max(1, int(round(H / fy))), max(1, int(round(W / s)))
        assert np.array_equal(op.resize_cubic_f32(d, dw, dh).view(np.uint32),
                              cv2.resize(d, (dw, dh), interpolation=cv2.INTER_CUBIC).view(np.uint32)), (H, W, s)
        assert np.array_equal(op.resize_nearest(d, dw, dh), cv2.resize(d, (dw, dh), interpolation=cv2.INTER_NEAREST))
