"""Pins oracle/pipeline.py's triangulate() -- the restatement the GPU triangulation kernel is compared with -- against the
REFERENCE'S OWN CODE: tests/golden/triang_golden.npz was produced by oracle/_ref/triang_ref, i.e. the reference's
`size_t triangulate( StereoMatchEnv& )` and `StereoMatchEnv::unrectify` (src/wass_stereo/wass_stereo.cpp:299-324, 1039-1386)
cut out of its source at build time and compiled with its PovMesh.cpp / triangulate.hpp against the header shim
(tests/golden/make_triang_golden.py, oracle/build_ref.sh, oracle/cut_triangulate.awk).  Runs on the CPU.

Which grid slots hold a point (every gate: disparity, rectified range, image border, bounding box, masks, burned areas,
minimum angle, distance limits): exact.  Point coordinates: 1e-12 relative (same formulas, numpy vs cv::Matx operation
grouping).  Grey values: exact."""
import os

import numpy as np
import pytest

from helpers import GOLDEN
from oracle import pipeline as op

Z = np.load(os.path.join(GOLDEN, "triang_golden.npz"))


def _cfg(name):
    kv = {}
    for line in bytes(Z[name + "/config"]).decode().splitlines():
        if "=" in line:
            k, v = line.split("=", 1)
            kv[k.strip()] = v.strip().strip('"')
    return kv


@pytest.mark.parametrize("name", [str(n) for n in Z["names"]])
def test_triangulate_matches_reference(name):
    g = lambda k: Z[name + "/" + k]
    kv = _cfg(name)
    calib = dict(K0=g("K0"), K1=g("K1"), R=g("R"), T=g("T"), R1=g("R1"), R2=g("R2"), P1=g("P1"), P2=g("P2"),
                 roi_left=tuple(int(v) for v in g("roiL")), roi_right=tuple(int(v) for v in g("roiR")))
    if kv.get("USE_CUSTOM_STEREORECTIFY") == "true":
        calib["HLi"], calib["HRi"] = g("HLi"), g("HRi")
    bbox = None
    if all(float(kv.get(k, -1)) >= 0 for k in ("TRIANG_BBOX_TOP", "TRIANG_BBOX_LEFT", "TRIANG_BBOX_BOTTOM", "TRIANG_BBOX_RIGHT")):
        bbox = (float(kv["TRIANG_BBOX_LEFT"]), float(kv["TRIANG_BBOX_TOP"]), float(kv["TRIANG_BBOX_RIGHT"]), float(kv["TRIANG_BBOX_BOTTOM"]))
    comp, camdist = g("scal")
    r = op.triangulate(g("disp"), calib, g("left"), g("right"),
                       left_mask=g("lmask") if "LEFT_MASK_IMAGE" in kv else None,
                       right_mask=g("rmask") if "RIGHT_MASK_IMAGE" in kv else None,
                       min_angle=float(kv.get("TRIANG_MIN_ANGLE", 20.0)), bbox=bbox,
                       discard_burned=kv.get("DISCARD_BURNED_AREAS", "true") == "true",
                       disparity_compensation=float(comp), dense_scale=float(kv.get("DENSE_SCALE", 1.0)), cam_distance=float(camdist))
    ref_valid = g("valid").astype(bool)
    assert r["n"] == int(g("n")[0]) == int(ref_valid.sum())
    assert np.array_equal(r["valid"], ref_valid)
    a, b = r["p3d"][ref_valid], g("xyz")[ref_valid]
    assert np.allclose(a, b, rtol=1e-12, atol=1e-12), float(np.abs(a - b).max())
    assert np.array_equal(r["color"][ref_valid], g("grey")[ref_valid])
