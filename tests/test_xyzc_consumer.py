"""Consumer side of mesh_cam.xyzC: load_camera_mesh + align_on_sea_plane of
gridding/wassgridsurface/wass_utils.py.  Golden vectors come from the reference's own functions
(tests/golden/make_xyzc_golden.py); the oracle restatement and the device op are both held to them."""
import os
import numpy as np
import pytest
from helpers import GOLDEN

G = np.load(os.path.join(GOLDEN, "xyzc_golden.npz"))
TOL = dict(rtol=1e-12, atol=1e-12)   # float64 3x3 products: only the summation order / FMA contraction may differ


def test_oracle_reader_matches_reference_reader():
    from oracle import pipeline as op
    buf = G["xyzc"].tobytes()
    mesh = op.load_camera_mesh_bytes(buf)
    assert mesh.shape == G["mesh_cam"].shape
    assert np.allclose(mesh, G["mesh_cam"], **TOL)
    aligned = op.align_on_sea_plane(mesh, G["mean_plane"], float(G["baseline"]))
    assert np.allclose(aligned, G["aligned"], **TOL)
    # the writer side of the oracle agrees with the reader to the quantisation step (one u16 LSB per axis)
    P = G["p3d"][G["valid"]]
    lim = np.frombuffer(buf[4:52], "<f8")
    lsb = 1.0 / lim[:3]
    assert np.all(np.abs(mesh.T - P).max(axis=0) <= 2.0 * np.linalg.norm(lsb))


@pytest.mark.gpu
def test_device_decode_align_matches_reference():
    from wass_b200 import capi
    h = capi.Handle(0)
    try:
        out = h.xyzc_decode_align(G["xyzc"].tobytes(), G["mean_plane"], float(G["baseline"]))
        assert out.shape == G["aligned"].shape
        assert np.allclose(out, G["aligned"], **TOL)
        # empty mesh
        hdr = bytearray(G["xyzc"].tobytes()[:148]); hdr[0:4] = (0).to_bytes(4, "little")
        assert h.xyzc_decode_align(bytes(hdr), G["mean_plane"]).shape == (3, 0)
        # truncated file is an error, not a read past the end
        with pytest.raises(capi.WsgError):
            h.xyzc_decode_align(G["xyzc"].tobytes()[:1000], G["mean_plane"])
    finally:
        h.close()


@pytest.mark.gpu
def test_device_mesh_to_aligned_points_without_the_file():
    """wsg_mesh_aligned_points == decode(export(mesh)) bit for bit, and equals the reference reader on the same bytes."""
    from wass_b200 import capi
    h = capi.Handle(0)
    try:
        h.mesh_upload(G["valid"], G["p3d"])
        plane, mean_plane, b = G["plane"], G["mean_plane"], float(G["baseline"])
        direct = h.mesh_aligned_points(plane, mean_plane, b)
        via_file = h.xyzc_decode_align(h.mesh_export_xyzc(plane), mean_plane, b)
        assert np.array_equal(direct, via_file)
        # the device writer may differ from the oracle writer by one u16 LSB (fp64 rounding of the plane rotation)
        lim = np.frombuffer(G["xyzc"].tobytes()[4:28], "<f8")
        assert direct.shape == G["aligned"].shape
        assert np.abs(direct - G["aligned"]).max() <= 2.5 * b * np.linalg.norm(1.0 / lim)
    finally:
        h.close()
