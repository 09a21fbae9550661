"""CPU-side checks of the boundary: libwassgpu.so loads, exports every symbol include/wassgpu.h declares,
and its host-only entry points work.  No device compute here."""
import ctypes
import os
import re
import numpy as np
import pytest
from helpers import ROOT


def _declared():
    txt = open(os.path.join(ROOT, "include", "wassgpu.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(wsg_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from wass_b200 import capi
    lib = capi.load()
    names = _declared()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_no_cuda_device_is_a_loud_error():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from wass_b200 import capi
    with pytest.raises(capi.WsgError):
        capi.Handle(0)


def test_host_only_entry_points():
    from wass_b200 import capi
    from oracle import pipeline as op
    plane = np.array([0.1, -0.5, 0.8]); plane = np.append(plane / np.linalg.norm(plane), -3.0)
    R, T, Ri, Ti = capi.rt_from_plane(plane)
    Ro, To, Rio, Tio = op.rt_from_plane(*plane)
    assert np.array_equal(R, Ro) and np.array_equal(T, To) and np.array_equal(Ri, Rio) and np.array_equal(Ti, Tio)
    mean, acc = capi.plane_mean([[1, 2, 3, 4], [np.nan] * 4, [3, 2, 1, 0]])
    assert np.array_equal(mean, [2, 2, 2, 2]) and acc[4] == 2
    assert np.isnan(capi.plane_mean([[np.nan] * 4])[0]).all()
    # libc rand() driven draw == the oracle's restatement of PovMesh.cpp:678-692
    libc = ctypes.CDLL("libc.so.6")
    libc.srand(777)
    a = capi.ransac_draw(120, 90, 40)
    b = op.ransac_draw_triples(op.LibcRand(777), 120, 90, 40)
    assert np.array_equal(a, b)


def test_defaults_match_reference_config_defaults():
    from wass_b200 import capi
    d = capi.dense_params()
    assert (d.MIN_DISPARITY, d.MAX_DISPARITY, d.WINSIZE, d.DENSE_P1_MULT, d.DENSE_P2_MULT) == (1, 640, 13, 2, 64)
    assert (d.DENSE_UNIQUENESS_RATIO, d.DENSE_DISP12MAXDIFF, d.DENSE_PREFILTER_CAP, d.DENSE_SPECKLE_WINDOW_SIZE) == (1, -1, 60, -70)
    assert (d.DISP_DILATE_STEPS, d.DISP_EROSION_STEPS, d.mode) == (1, 2, 0)
    t = capi.tri_params()
    assert t.TRIANG_MIN_ANGLE == 20.0 and t.DISCARD_BURNED_AREAS == 1 and t.cam_distance == 1.0
    r = capi.refine_params()
    assert r.PLANE_REFINEMENT_MAX_DISTANCE == 70.0 and r.PLANE_WEIGHT_PROPORTIONAL_TO_DISTANCE == 1
