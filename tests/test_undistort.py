"""cv::undistort as wass_prepare applies it (src/wass_prepare/wass_prepare.cpp:268) on the device,
against golden vectors from cv2.undistort (tests/golden/make_undistort_golden.py).  u8, bit-exact."""
import os
import numpy as np
import pytest
from helpers import GOLDEN

G = np.load(os.path.join(GOLDEN, "undistort_golden.npz"))
N = int(G["n"])


def test_golden_file_is_complete():
    for i in range(N):
        assert G["img_%d" % i].shape == G["out_%d" % i].shape and G["K_%d" % i].shape == (3, 3)


@pytest.mark.gpu
@pytest.mark.parametrize("i", range(N))
def test_device_undistort_matches_cv2(i):
    from wass_b200 import capi
    h = capi.Handle(0)
    try:
        out = h.undistort_image(G["img_%d" % i], G["K_%d" % i], G["dist_%d" % i])
        ref = G["out_%d" % i]
        nbad = int((out != ref).sum())
        assert nbad == 0, "%d of %d pixels differ (max |diff| %d)" % (nbad, ref.size, int(np.abs(out.astype(int) - ref).max()))
    finally:
        h.close()
