"""Optional refinement of the float ROI disparity (wass_stereo.cpp:941-986: MEDIAN_FILTER_WSIZE,
DENSE_DISPARITY_BIGGEST_COMPONENT_THRESHOLD; both off at the reference defaults).  Golden vectors: cv2.medianBlur /
cv2.Sobel / cv2.connectedComponentsWithStats (tests/golden/make_refine_golden.py).  float32, bit-exact."""
import os
import numpy as np
import pytest
from helpers import GOLDEN

G = np.load(os.path.join(GOLDEN, "refine_golden.npz"))
N = int(G["n"])


@pytest.mark.parametrize("i", range(N))
def test_oracle_matches_cv2(i):
    from oracle import pipeline as op
    m, t = (int(v) for v in G["par_%d" % i])
    out = op.refine_disparity(G["in_%d" % i], m, t)
    assert np.array_equal(out, G["out_%d" % i])


@pytest.mark.gpu
@pytest.mark.parametrize("i", range(N))
def test_device_matches_cv2(i):
    from wass_b200 import capi
    h = capi.Handle(0)
    try:
        m, t = (int(v) for v in G["par_%d" % i])
        out = h.disparity_refine(G["in_%d" % i], m, t)
        assert np.array_equal(out, G["out_%d" % i])
    finally:
        h.close()


@pytest.mark.gpu
def test_unsupported_median_size_is_an_error():
    from wass_b200 import capi
    h = capi.Handle(0)
    try:
        with pytest.raises(capi.WsgError):
            h.disparity_refine(np.ones((8, 8), np.float32), 7, 0)
    finally:
        h.close()
