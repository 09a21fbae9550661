"""Parity of the CUDA dense matcher (through the C ABI) against the CPU oracle and the cv2 golden vectors.
Integer outputs: bit-exact, every pixel."""
import numpy as np
import pytest
from helpers import load_sgbm_golden

pytestmark = pytest.mark.gpu

CASES = load_sgbm_golden()


@pytest.fixture(scope="module")
def handle():
    from wass_b200 import capi
    h = capi.Handle(0)
    yield h
    h.close()


def _supported(p):
    return True


IMPLS = [0, 1, 2]   # capi.AGG_*: three device decompositions of the aggregation, one result


@pytest.fixture(params=IMPLS, ids=["per_direction", "sweeps", "sweeps_wta"])
def impl(request, handle):
    handle.sgbm_set_impl(request.param)
    yield request.param
    handle.sgbm_set_impl(2)


@pytest.mark.parametrize("idx", range(len(CASES)))
def test_gpu_matches_cv2_golden(handle, impl, idx):
    img1, img2, p, disp = CASES[idx]
    if not _supported(p):
        pytest.skip("speckle filter not implemented on GPU yet")
    out = handle.sgbm_compute(img1, img2, p)
    nbad = int((out != disp).sum())
    assert nbad == 0, "%d / %d pixels differ from cv2" % (nbad, disp.size)
    want = impl if p["numDisparities"] <= 512 else 0
    assert handle.sgbm_stats()["agg_impl"] == want


@pytest.mark.parametrize("idx", [0, 3, 9, 12])
@pytest.mark.parametrize("which", [0, 1], ids=["per_direction", "sweeps"])
def test_gpu_volumes_match_oracle(handle, which, idx):
    from oracle import sgbm
    img1, img2, p, _ = CASES[idx]
    ref = sgbm.compute(img1, img2, p, want_volumes=True)
    handle.sgbm_set_impl(which)
    try:
        handle.sgbm_compute(img1, img2, p)
        H, W1, D = ref["C"].shape
        C, S = handle.sgbm_debug_volumes(H, W1, D)
    finally:
        handle.sgbm_set_impl(2)
    assert np.array_equal(C, ref["C"]), "cost volume differs"
    assert np.array_equal(S, ref["S"]), "aggregated volume differs"
    st = handle.sgbm_stats()
    assert st["max_cost"] == ref["maxC"]
    assert st["out_of_domain"] == int(ref["maxC"] + max(p["P2"], p["P1"] + 1) > 32767)


@pytest.mark.parametrize("W,H,D,mode", [(640, 480, 64, 0), (640, 480, 64, 1), (500, 120, 256, 1),
                                        (300, 64, 512, 1), (260, 48, 640, 0), (333, 77, 80, 1)])
def test_gpu_matches_oracle_wass_defaults(handle, impl, W, H, D, mode):
    from oracle import sgbm
    from wass_b200 import synth
    r, l, _ = synth.make_pair(W, H, D, seed=W + D)
    i1, i2 = synth.pad_for_sgbm(r, l, D)
    p = sgbm.wass_params(D, mode=mode)
    ref = sgbm.compute(i1, i2, p)
    out = handle.sgbm_compute(i1, i2, p)
    assert ref["maxC"] + p["P2"] <= 32767
    assert np.array_equal(out, ref["disp"])


def test_fused_wta_does_not_expose_S(handle):
    """AGG_SWEEPS_WTA consumes S inside the last sweep; asking for it is a state error, not stale data."""
    from wass_b200 import capi
    img1, img2, p, _ = CASES[0]
    handle.sgbm_set_impl(capi.AGG_SWEEPS_WTA)
    handle.sgbm_compute(img1, img2, p)
    H, W = img1.shape
    with pytest.raises(capi.WsgError) as e:
        handle.sgbm_debug_volumes(H, 8, p["numDisparities"])
    assert e.value.code == -5


@pytest.mark.parametrize("mode", [0, 1])
def test_sweeps_many_bands_and_reuse(handle, mode):
    """Tall image (many 8-row bands, H not a multiple of 8), then a different geometry on the same handle:
    the band hand-off buffer and its epoch tags must survive reuse."""
    from oracle import sgbm
    from wass_b200 import capi, synth
    handle.sgbm_set_impl(capi.AGG_SWEEPS_WTA)
    for (W, H, D) in [(96, 331, 64), (96, 331, 64), (200, 75, 160), (96, 331, 64)]:
        r, l, _ = synth.make_pair(W, H, D, seed=H + D)
        i1, i2 = synth.pad_for_sgbm(r, l, D)
        p = sgbm.wass_params(D, mode=mode)
        ref = sgbm.compute(i1, i2, p)
        for _ in range(2):
            out = handle.sgbm_compute(i1, i2, p)
            assert np.array_equal(out, ref["disp"])


@pytest.mark.parametrize("workers", [1, 3, 74, 1000])
def test_sweep_worker_cap_does_not_change_results(handle, workers):
    """wsg_sgbm_set_sweep_workers: however few SMs take the bands of a sweep (one worker walks all of them in ticket order),
    the disparity is the oracle's; a cap above the SM count means all of them."""
    from oracle import sgbm
    from wass_b200 import capi, synth
    handle.sgbm_set_impl(capi.AGG_SWEEPS_WTA)
    try:
        handle.sgbm_set_sweep_workers(workers)
        for mode, (W, H, D) in ((1, (96, 331, 64)), (0, (200, 75, 160))):
            r, l, _ = synth.make_pair(W, H, D, seed=H + D + 1)
            i1, i2 = synth.pad_for_sgbm(r, l, D)
            p = sgbm.wass_params(D, mode=mode)
            assert np.array_equal(handle.sgbm_compute(i1, i2, p), sgbm.compute(i1, i2, p)["disp"])
        with pytest.raises(capi.WsgError):
            handle.sgbm_set_sweep_workers(-1)
    finally:
        handle.sgbm_set_sweep_workers(0)


@pytest.mark.parametrize("uniq", [40, 99, 100, 150])
def test_uniqueness_ratio_extremes(handle, impl, uniq):
    """Large uniquenessRatio, and >= 100 (100-uniq <= 0: the in-sweep WTA hands over to wta_kernel).  The oracle agrees
    with cv2 4.13 on these (checked in tests/test_oracle_golden.py::test_uniqueness_extremes_vs_cv2 when cv2 is there)."""
    from oracle import sgbm
    from wass_b200 import synth
    r, l, _ = synth.make_pair(180, 40, 64, seed=uniq)
    i1, i2 = synth.pad_for_sgbm(r, l, 64)
    p = sgbm.wass_params(64, mode=1)
    p["uniquenessRatio"] = uniq
    ref = sgbm.compute(i1, i2, p)
    out = handle.sgbm_compute(i1, i2, p)
    assert np.array_equal(out, ref["disp"])


def test_too_narrow_image_is_an_error(handle):
    from wass_b200 import capi
    a = np.zeros((8, 20), np.uint8)
    p = dict(minDisparity=1, numDisparities=32, blockSize=5, P1=200, P2=800, disp12MaxDiff=1,
             preFilterCap=60, uniquenessRatio=5, speckleWindowSize=0, speckleRange=0, mode=0)
    with pytest.raises(capi.WsgError) as e:
        handle.sgbm_compute(a, a, p)
    assert e.value.code == -4


@pytest.mark.parametrize("n", [1, 2, 3, 5])
@pytest.mark.parametrize("mode", [0, 1])
def test_batch_matches_oracle(handle, impl, n, mode):
    """wsg_sgbm_compute_batch: n different frames in one call (one sweep launch walks the bands of all frames interleaved);
    every frame must be the oracle's.  H is not a multiple of the band height; several bands per frame."""
    from oracle import sgbm
    from wass_b200 import synth
    W, H, D = 120, 53, 64
    frames = [synth.pad_for_sgbm(*synth.make_pair(W, H, D, seed=100 * n + f)[:2], D) for f in range(n)]
    p = sgbm.wass_params(D, mode=mode)
    outs = handle.sgbm_compute_batch([f[0] for f in frames], [f[1] for f in frames], p)
    for f in range(n):
        assert np.array_equal(outs[f], sgbm.compute(frames[f][0], frames[f][1], p)["disp"]), "frame %d of %d" % (f, n)
    st = handle.sgbm_stats()
    assert st["agg_impl"] == impl and st["out_of_domain"] == 0


def test_batch_then_single_then_other_batch_size(handle):
    """The hand-off buffer and its epoch tags across changing batch sizes on one handle."""
    from oracle import sgbm
    from wass_b200 import capi, synth
    handle.sgbm_set_impl(capi.AGG_SWEEPS_WTA)
    W, H, D = 96, 100, 64
    p = sgbm.wass_params(D, mode=1)
    frames = [synth.pad_for_sgbm(*synth.make_pair(W, H, D, seed=900 + f)[:2], D) for f in range(4)]
    refs = [sgbm.compute(a, b, p)["disp"] for a, b in frames]
    for n in (4, 1, 3, 3, 2, 4):
        outs = handle.sgbm_compute_batch([f[0] for f in frames[:n]], [f[1] for f in frames[:n]], p)
        for f in range(n):
            assert np.array_equal(outs[f], refs[f]), "batch of %d, frame %d" % (n, f)


def test_async_batches_match_the_synchronous_call(handle):
    """wsg_sgbm_batch_submit / _wait: two batches in flight on one handle (copies on their own streams); results are those
    of wsg_sgbm_compute_batch, slot misuse is a state error."""
    import torch
    from oracle import sgbm
    from wass_b200 import capi, synth
    handle.sgbm_set_impl(capi.AGG_SWEEPS_WTA)
    W, H, D, n = 150, 61, 64, 3
    p = sgbm.wass_params(D, mode=1)
    sets = []
    for k in range(3):
        fr = [synth.pad_for_sgbm(*synth.make_pair(W, H, D, seed=700 + 10 * k + f)[:2], D) for f in range(n)]
        sets.append(fr)
    refs = [handle.sgbm_compute_batch([f[0] for f in fr], [f[1] for f in fr], p) for fr in sets]
    Hh, Wp = sets[0][0][0].shape
    pins = [([torch.from_numpy(f[0]).pin_memory() for f in fr], [torch.from_numpy(f[1]).pin_memory() for f in fr],
             [torch.empty((Hh, Wp), dtype=torch.int16).pin_memory() for _ in fr]) for fr in sets]

    def submit(slot, k):
        a, b, d = pins[k]
        handle.sgbm_batch_submit(slot, n, [t.data_ptr() for t in a], [t.data_ptr() for t in b], Hh, Wp, Wp, p, [t.data_ptr() for t in d])

    with pytest.raises(capi.WsgError) as e:
        handle.sgbm_batch_wait(0)
    assert e.value.code == -5
    submit(0, 0)
    submit(1, 1)
    with pytest.raises(capi.WsgError) as e:
        submit(0, 2)                      # slot 0 has not been waited for
    assert e.value.code == -5
    handle.sgbm_batch_wait(0)
    submit(0, 2)
    handle.sgbm_batch_wait(1)
    handle.sgbm_batch_wait(0)
    for k in range(3):
        for f in range(n):
            assert np.array_equal(pins[k][2][f].numpy(), refs[k][f]), "set %d frame %d" % (k, f)


def test_rows_per_band_do_not_change_results(handle):
    """A launch may use fewer rows per band than the worker has row warps (sweep_rows_per_band picks them by batch size);
    WSG_SWEEP_ROWS is read once per process, so this test drives the choice through the batch size instead."""
    from oracle import sgbm
    from wass_b200 import capi, synth
    handle.sgbm_set_impl(capi.AGG_SWEEPS_WTA)
    W, H, D = 100, 230, 64                 # 230 rows: 16 bands of 15, 17 of 14, 18 of 13 -- the wave model moves with n
    p = sgbm.wass_params(D, mode=1)
    frames = [synth.pad_for_sgbm(*synth.make_pair(W, H, D, seed=40 + f)[:2], D) for f in range(9)]
    refs = [sgbm.compute(a, b, p)["disp"] for a, b in frames[:3]]
    for n in (1, 2, 3, 5, 9):
        outs = handle.sgbm_compute_batch([f[0] for f in frames[:n]], [f[1] for f in frames[:n]], p)
        for f in range(min(n, 3)):
            assert np.array_equal(outs[f], refs[f]), "batch of %d, frame %d" % (n, f)


def test_batch_argument_errors(handle):
    from wass_b200 import capi
    a = np.zeros((16, 200), np.uint8)
    p = dict(minDisparity=1, numDisparities=32, blockSize=5, P1=200, P2=800, disp12MaxDiff=1,
             preFilterCap=60, uniquenessRatio=5, speckleWindowSize=0, speckleRange=0, mode=0)
    with pytest.raises(ValueError):
        handle.sgbm_compute_batch([], [], p)
    with pytest.raises(capi.WsgError) as e:
        handle.sgbm_compute_batch([a] * (capi.MAX_BATCH + 1), [a] * (capi.MAX_BATCH + 1), p)
    assert e.value.code == -1


EDGE = [
    # (W, H, D, minD, block, mode)   -- W is the image width INCLUDING the maxD columns cv2 leaves invalid
    (40, 1, 16, 0, 3, 1),        # a single row
    (40, 2, 16, 1, 5, 0),
    (60, 7, 32, 1, 7, 1),        # fewer rows than one 8-row band
    (60, 9, 32, 1, 7, 1),        # one band and one row
    (23, 17, 16, 0, 1, 1),       # blockSize 1, 7 matched columns
    (36, 20, 16, 3, 9, 0),
    (30, 24, 16, -5, 5, 1),      # negative minDisparity
    (130, 33, 96, 2, 25, 1),     # window 25: the narrow-tile cost kernel (the wide one stops at 17)
    (130, 33, 96, 2, 19, 0),
    (1400, 12, 1280, 0, 5, 1),   # 1280 disparities: more than the fused sweeps take, per-direction launches
]


@pytest.mark.parametrize("W,H,D,minD,block,mode", EDGE)
def test_edge_shapes_match_oracle(handle, impl, W, H, D, minD, block, mode):
    """Ragged and extreme shapes: single rows, partial bands, a handful of matched columns, negative minDisparity,
    the largest window and disparity count the API accepts."""
    from oracle import sgbm
    rng = np.random.default_rng(W * 131 + H)
    coarse = rng.integers(90, 170, (H // 3 + 2, W // 3 + 2)).astype(np.float32)
    img = np.kron(coarse, np.ones((3, 3), np.float32))[:H, :W]
    img1 = np.clip(img + rng.normal(0, 2, img.shape), 0, 255).astype(np.uint8)
    img2 = np.clip(np.roll(img, -3, axis=1) + rng.normal(0, 2, img.shape), 0, 255).astype(np.uint8)
    p = dict(minDisparity=minD, numDisparities=D, blockSize=block, P1=min(8 * block * block, 1200),
             P2=min(32 * block * block, 5000), disp12MaxDiff=1, preFilterCap=31, uniquenessRatio=5, speckleWindowSize=0,
             speckleRange=0, mode=mode)
    ref = sgbm.compute(img1, img2, p)
    assert ref["maxC"] + p["P2"] <= 32767
    out = handle.sgbm_compute(img1, img2, p)
    assert np.array_equal(out, ref["disp"])
