"""Pins the CPU oracle (oracle/sgbm_oracle.c) against golden vectors produced by the real
cv2.StereoSGBM (tests/golden/make_golden.py) -- bit-exact on every pixel."""
import numpy as np
import pytest
from helpers import load_sgbm_golden
from oracle import sgbm

CASES = load_sgbm_golden()


@pytest.mark.parametrize("idx", range(len(CASES)))
def test_oracle_matches_cv2_golden(idx):
    img1, img2, p, disp = CASES[idx]
    out = sgbm.compute(img1, img2, p)
    assert out["disp"].dtype == np.int16 and out["disp"].shape == disp.shape
    assert np.array_equal(out["disp"], disp)


def test_oracle_matches_live_cv2_if_present():
    cv2 = pytest.importorskip("cv2")
    from wass_b200 import synth
    r, l, _ = synth.make_pair(150, 40, 48, seed=11)
    i1, i2 = synth.pad_for_sgbm(r, l, 48)
    for mode in (0, 1):
        p = sgbm.wass_params(48, mode=mode)
        m = cv2.StereoSGBM_create(p["minDisparity"], 48, 13, p["P1"], p["P2"])
        m.setUniquenessRatio(1); m.setDisp12MaxDiff(-1); m.setPreFilterCap(60)
        m.setSpeckleRange(16); m.setSpeckleWindowSize(-70)
        m.setMode(cv2.STEREO_SGBM_MODE_HH if mode else cv2.STEREO_SGBM_MODE_SGBM)
        assert np.array_equal(m.compute(i1, i2), sgbm.compute(i1, i2, p)["disp"])


def test_oracle_volumes_shape_and_domain():
    img1, img2, p, _ = CASES[0]
    out = sgbm.compute(img1, img2, p, want_volumes=True, want_raw=True)
    H, W = img1.shape
    W1 = W - (p["minDisparity"] + p["numDisparities"])
    assert out["C"].shape == (H, W1, p["numDisparities"])
    assert out["C"].max() == out["maxC"]
    assert out["maxC"] + p["P2"] <= 32767  # fixtures stay inside the verified domain
    assert (out["S"] >= out["C"]).all()


def test_uniqueness_extremes_vs_cv2():
    """uniquenessRatio in {1, 40, 99, 100, 150}: the oracle equals cv2 where cv2 is importable (it is not on the GPU box).
    uniquenessRatio == 0 is OUTSIDE the verified domain: cv2 4.13 then also drops pixels whose minimum is tied by a
    disparity further than one step away (saturated S), which the published inequality S(d)*(100-u) < minS*100 never
    does; the reference's default is 1 (wass_stereo.cpp:755)."""
    cv2 = pytest.importorskip("cv2")
    from oracle import sgbm
    from wass_b200 import synth
    for uniq in (1, 40, 99, 100, 150):
        r, l, _ = synth.make_pair(180, 40, 64, seed=uniq)
        i1, i2 = synth.pad_for_sgbm(r, l, 64)
        p = sgbm.wass_params(64, mode=1)
        p["uniquenessRatio"] = uniq
        m = cv2.StereoSGBM_create(p["minDisparity"], p["numDisparities"], p["blockSize"], p["P1"], p["P2"])
        m.setUniquenessRatio(uniq); m.setDisp12MaxDiff(p["disp12MaxDiff"]); m.setPreFilterCap(p["preFilterCap"])
        m.setSpeckleRange(p["speckleRange"]); m.setSpeckleWindowSize(p["speckleWindowSize"]); m.setMode(1)
        assert np.array_equal(m.compute(i1, i2), sgbm.compute(i1, i2, p)["disp"]), uniq
