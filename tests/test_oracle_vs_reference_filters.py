"""Pins oracle/pipeline.py's restatement of the reference's disparity clean-up (matrix_dilate_zero incl. its one-column shift,
matrix_erode_zero, clean_and_convert_disparity: src/wass_stereo/wass_stereo.cpp:617-733) against the REFERENCE'S OWN CODE:
tests/golden/filters_golden.npz was produced by oracle/_ref/filters_ref, the three functions cut out of the reference source at
build time and compiled against the header shim (tests/golden/make_filters_golden.py, oracle/build_ref.sh).  Bit for bit, on CPU."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from helpers import GOLDEN, ROOT
from oracle import pipeline as op

Z = np.load(os.path.join(GOLDEN, "filters_golden.npz"))


@pytest.mark.parametrize("name", [str(n) for n in Z["names"]])
def test_dilate_erode_match_reference(name):
    src = Z[name + "/src"]
    assert np.array_equal(op.matrix_dilate_zero(src).view(np.uint32), Z[name + "/dilate"].view(np.uint32))
    assert np.array_equal(op.matrix_erode_zero(src).view(np.uint32), Z[name + "/erode"].view(np.uint32))
    chain = op.matrix_erode_zero(op.matrix_dilate_zero(op.matrix_dilate_zero(src)))      # DISP_DILATE_STEPS=2, DISP_EROSION_STEPS=1
    assert np.array_equal(chain.view(np.uint32), Z[name + "/dilate2_erode"].view(np.uint32))


@pytest.mark.parametrize("name", [str(n) for n in Z["cnames"]])
def test_clean_and_convert_matches_reference(name):
    mind, nd, off, sc = Z[name + "/args"]
    got = op.clean_and_convert_disparity(Z[name + "/src"], int(mind), int(nd), int(off), float(sc))
    assert np.array_equal(got.view(np.uint32), Z[name + "/clean"].view(np.uint32))


def test_live_binary_agrees_with_golden_when_the_reference_is_here():
    """In this container the reference is present: rebuild the binary and re-run one case live."""
    if not os.path.isdir("/root/reference"):
        pytest.skip("no reference tree (GPU box)")
    subprocess.run(["bash", os.path.join(ROOT, "oracle", "build_ref.sh")], check=True, capture_output=True)
    exe = os.path.join(ROOT, "oracle", "_ref", "filters_ref")
    src = Z["f0/src"]
    with tempfile.TemporaryDirectory() as td:
        fi, fo = os.path.join(td, "i"), os.path.join(td, "o")
        src.tofile(fi)
        subprocess.run([exe, "dilate", fi, str(src.shape[0]), str(src.shape[1]), fo], check=True)
        assert np.array_equal(np.fromfile(fo, np.float32).reshape(src.shape).view(np.uint32), Z["f0/dilate"].view(np.uint32))
