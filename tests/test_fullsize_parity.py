"""Parity at the sizes BASELINE.json names, against the real cv2.StereoSGBM (wass_stereo.cpp:837 calls it): the
benchmark frames of bench.py (2448x2048x256 MODE_HH, seeds 0..2), configs[1] in both modes and configs[3]
(4096x3000x512).  Bit-exact, every pixel.

The cv2 answers are pinned as SHA-256 in tests/golden/fullsize_hashes.json (generated here in the build container by
tests/golden/make_fullsize_hashes.py from cv2 itself).  If the seeded inputs hash differently on the test machine
(another numpy / cv2 resize build) the test runs cv2 there instead, so it never compares against the wrong frame."""
import json
import os

import numpy as np
import pytest

from helpers import GOLDEN

pytestmark = pytest.mark.gpu

with open(os.path.join(GOLDEN, "fullsize_hashes.json")) as f:
    HASHES = json.load(f)


def _sha(*arrays):
    import hashlib
    h = hashlib.sha256()
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(str(a.shape).encode() + str(a.dtype).encode())
        h.update(a.tobytes())
    return h.hexdigest()


@pytest.fixture(scope="module")
def handle():
    from wass_b200 import capi
    h = capi.Handle(0)
    yield h
    h.close()


def _check(name, out, i1, i2, p):
    c = HASHES["cases"][name]
    if _sha(i1, i2) == c["inputs_sha256"]:
        assert _sha(out) == c["disp_sha256"], "%s: GPU disparity differs from cv2 %s (pinned hash)" % (name, HASHES["cv2"])
        return "pinned"
    cv2 = pytest.importorskip("cv2")       # inputs differ on this machine: the live cv2 is the oracle
    cv2.setNumThreads(1)
    m = cv2.StereoSGBM_create(p["minDisparity"], p["numDisparities"], p["blockSize"], p["P1"], p["P2"])
    m.setUniquenessRatio(p["uniquenessRatio"]); m.setDisp12MaxDiff(p["disp12MaxDiff"])
    m.setPreFilterCap(p["preFilterCap"]); m.setSpeckleRange(p["speckleRange"]); m.setSpeckleWindowSize(p["speckleWindowSize"])
    m.setMode(cv2.STEREO_SGBM_MODE_HH if p["mode"] == 1 else cv2.STEREO_SGBM_MODE_SGBM)
    ref = m.compute(i1, i2)
    nbad = int((out != ref).sum())
    assert nbad == 0, "%s: %d / %d pixels differ from the live cv2" % (name, nbad, ref.size)
    return "live"


@pytest.mark.parametrize("name", ["config2_hh", "config2_sgbm", "config4_hh"])
def test_full_size_matches_cv2(handle, name):
    """BASELINE configs[1] (both modes) and configs[3]: the product path (fused sweeps + fused WTA) against cv2, the
    per-direction decomposition against the product path, and the generator's ground truth as a sanity check."""
    from oracle import sgbm
    from wass_b200 import capi, synth
    c = HASHES["cases"][name]
    W, H, D, mode = c["W"], c["H"], c["D"], c["mode"]
    r, l, d_true = synth.make_pair(W, H, D, seed=c["seed"])
    i1, i2 = synth.pad_for_sgbm(r, l, D)
    p = sgbm.wass_params(D, mode=mode)
    handle.sgbm_set_impl(capi.AGG_SWEEPS_WTA)
    out = handle.sgbm_compute(i1, i2, p).copy()
    st = handle.sgbm_stats()
    assert st["agg_impl"] == capi.AGG_SWEEPS_WTA and st["out_of_domain"] == 0
    _check(name, out, i1, i2, p)
    try:
        for impl in (capi.AGG_SWEEPS, capi.AGG_PER_DIRECTION):
            handle.sgbm_set_impl(impl)
            assert np.array_equal(handle.sgbm_compute(i1, i2, p), out), "implementation %d differs" % impl
    finally:
        handle.sgbm_set_impl(capi.AGG_SWEEPS_WTA)
    disp = out[:, D:].astype(np.float32) / 16.0
    valid = disp > 1
    assert valid.mean() > 0.85
    err = np.abs(disp - d_true)[valid]
    assert np.median(err) < 0.25 and (err < 1.0).mean() > 0.97


def test_benchmark_batch_matches_cv2(handle):
    """The frames bench.py times (seeds 0..2, 2448x2048x256 MODE_HH) as ONE batch through wsg_sgbm_compute_batch."""
    from oracle import sgbm
    from wass_b200 import capi, synth
    names = ["bench_frame_seed0_hh", "bench_frame_seed1_hh", "bench_frame_seed2_hh"]
    frames = []
    for n in names:
        c = HASHES["cases"][n]
        r, l, _ = synth.make_pair(c["W"], c["H"], c["D"], seed=c["seed"])
        frames.append(synth.pad_for_sgbm(r, l, c["D"]))
    p = sgbm.wass_params(256, mode=1)
    handle.sgbm_set_impl(capi.AGG_SWEEPS_WTA)
    outs = handle.sgbm_compute_batch([f[0] for f in frames], [f[1] for f in frames], p)
    for n, o, (i1, i2) in zip(names, outs, frames):
        _check(n, o, i1, i2, p)
