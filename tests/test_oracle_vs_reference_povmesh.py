"""Pins oracle/pipeline.py -- the numpy restatement the GPU mesh kernels are compared with -- against the REFERENCE'S OWN
CODE: tests/golden/povmesh_golden.npz was produced by oracle/_ref/povmesh_ref, i.e. the unmodified
src/wass_stereo/PovMesh.cpp and src/wass_lib/triangulate.hpp of the reference compiled against the header shim in
oracle/shim/ (tests/golden/make_povmesh_golden.py, oracle/build_ref.sh).  Runs on the CPU.

z-gap percentile, component masks, crops: exact.  RANSAC with a fixed srand(): the same plane to the last bit (this is
what pins the order of the rand() calls).  Plane refinement: 1e-9 (3x3 SVD by two different solvers).  .xyzC: header
fields 1e-12, quantised points identical or +-1 LSB where the SVD's last bits move a point across a step."""
import os
import struct

import numpy as np
import pytest

from helpers import GOLDEN
from oracle import pipeline as op

Z = np.load(os.path.join(GOLDEN, "povmesh_golden.npz"))
NAMES = [str(n) for n in Z["names"]]


def _case(name):
    valid = Z[name + "/valid"].astype(bool)
    p3d = Z[name + "/p3d"]
    a = Z[name + "/args"]
    cfg = bytes(Z[name + "/config"]).decode()
    kw = dict(weight_by_distance="PLANE_WEIGHT_PROPORTIONAL_TO_DISTANCE=false" not in cfg,
              central_third="PLANE_USE_CENTRAL_THIRD_ONLY=true" in cfg, max_distance=70.0)
    for line in cfg.splitlines():
        if line.startswith("PLANE_REFINEMENT_MAX_DISTANCE="):
            kw["max_distance"] = float(line.split("=")[1])
    return valid, p3d, dict(seed=int(a[0]), rounds=int(a[1]), thr=a[2], zpct=a[3], maxd=a[4], bounds=a[5:9]), kw


@pytest.mark.parametrize("name", NAMES)
def test_mesh_stage_matches_reference(name):
    valid, p3d, a, kw = _case(name)
    H, W = valid.shape
    zgap = op.zgap_percentile(valid, p3d[..., 2], a["zpct"])
    assert zgap == Z[name + "/zgap"][0]                                    # PovMesh.cpp:888-926
    comp = op.biggest_component(valid, p3d[..., 2], zgap)
    assert np.array_equal(comp, Z[name + "/mask_component"].astype(bool))  # PovMesh.cpp:929-987
    triples = op.ransac_draw_triples(op.LibcRand(a["seed"]), W, H, a["rounds"])
    ok, plane, _ = op.ransac_find_plane(comp, p3d, triples, a["thr"])      # PovMesh.cpp:665-777
    assert ok == bool(Z[name + "/ransac_ok"][0])
    assert np.array_equal(plane, Z[name + "/plane_ransac"]), "RANSAC plane differs: rand() order or arithmetic"
    mask = comp
    if ok:
        mask = op.crop_plane(comp, p3d, plane, a["thr"])                   # PovMesh.cpp:780-815
        assert np.array_equal(mask, Z[name + "/mask_crop1"].astype(bool))
        plane, n_in = op.refine_plane(mask, p3d, *a["bounds"], **kw)       # PovMesh.cpp:581-660
        assert n_in == int(Z[name + "/n_refine_inliers"][0])
        assert np.allclose(plane, Z[name + "/plane_refined"], rtol=0, atol=1e-9)
        mask = op.crop_plane(mask, p3d, Z[name + "/plane_refined"], a["maxd"])
    assert np.array_equal(mask, Z[name + "/mask_final"].astype(bool))
    assert int(mask.sum()) == int(Z[name + "/n_final"][0])


@pytest.mark.parametrize("name", NAMES)
def test_xyzc_writer_matches_reference(name):
    """PovMesh::save_as_xyz_compressed (PovMesh.cpp:377-460) and save_as_xyz_binary (:346-375), byte level."""
    _, p3d, _, _ = _case(name)
    mask = Z[name + "/mask_final"].astype(bool)
    plane = Z[name + "/plane_refined"] if name + "/plane_refined" in Z else Z[name + "/plane_ransac"]
    ref = bytes(Z[name + "/mesh_cam_xyzC"])
    mine = op.xyz_compressed_bytes(mask, p3d, plane)
    assert len(mine) == len(ref)
    n = struct.unpack_from("<I", ref, 0)[0]
    assert struct.unpack_from("<I", mine, 0)[0] == n == int(mask.sum())
    hm, hr = np.array(struct.unpack_from("<18d", mine, 4)), np.array(struct.unpack_from("<18d", ref, 4))
    assert np.allclose(hm, hr, rtol=1e-12, atol=1e-12)
    qm = np.frombuffer(mine, "<u2", 3 * n, 148).astype(np.int64)
    qr = np.frombuffer(ref, "<u2", 3 * n, 148).astype(np.int64)
    assert np.abs(qm - qr).max() <= 1 and (qm != qr).mean() < 1e-3
    # the consumer-side reader of the reference decodes the reference's file to the input points (quantisation error)
    dec = op.xyz_compressed_decode(ref)
    step = 1.0 / hr[0:3].min()
    assert np.abs(dec - p3d[mask]).max() < 2.5 * step
    xb = bytes(Z[name + "/mesh_cam_xyzbin"])
    assert struct.unpack_from("<I", xb, 0)[0] == n
    assert np.array_equal(np.frombuffer(xb, "<f4", 3 * n, 4).reshape(n, 3), p3d[mask].astype(np.float32))


def test_ply_writer_layout():
    """PovMesh::save_as_ply_points (PovMesh.cpp:463-517): header text and 15-byte records, from the reference's own file."""
    name = NAMES[0]
    _, p3d, _, _ = _case(name)
    mask = Z[name + "/mask_final"].astype(bool)
    grey = Z[name + "/grey"]
    ply = bytes(Z[name + "/mesh_ply"])
    head, body = ply.split(b"end_header\n", 1)
    n = int(mask.sum())
    assert head.decode().split("\n")[:9] == ["ply", "format binary_little_endian 1.0", "element vertex %d" % n, "property float x",
                                             "property float y", "property float z", "property uchar red",
                                             "property uchar green", "property uchar blue"]
    rec = np.frombuffer(body, np.dtype([("p", "<f4", 3), ("c", "u1", 3)]))
    assert rec.shape[0] == n
    assert np.array_equal(rec["p"], p3d[mask].astype(np.float32))
    assert np.array_equal(rec["c"], np.repeat(grey[mask][:, None], 3, axis=1))


def test_component_tie_goes_to_first_found_column_major():
    valid = Z["tie/valid"].astype(bool)
    p3d = Z["tie/p3d"]
    zgap = op.zgap_percentile(valid, p3d[..., 2], 99.0)
    assert zgap == Z["tie/zgap"][0]
    comp = op.biggest_component(valid, p3d[..., 2], zgap)
    assert np.array_equal(comp, Z["tie/mask_component"].astype(bool))
    assert comp[5:7, 0:3].all() and not comp[1:3, 6:9].any()      # the block in column 0 wins, not the one in row 1


def test_triangulate_matches_reference():
    """triangulate(p, q, R, T), src/wass_lib/triangulate.hpp:26-72, 200 random rigs."""
    items, ref = Z["tri/items"], Z["tri/xyz"]
    for it, r in zip(items, ref):
        x = op.triangulate_point(it[0:2], it[2:4], it[4:13].reshape(3, 3), it[13:16])
        assert np.allclose(x, r, rtol=1e-12, atol=1e-12)


def test_rt_from_plane_matches_reference():
    for i, pl in enumerate(Z["rt/planes"]):
        R, T, Rinv, Tinv = op.rt_from_plane(*pl)
        assert np.array_equal(R.reshape(-1), Z["rt/%d/R" % i]) and np.array_equal(T, Z["rt/%d/T" % i])
        assert np.array_equal(Rinv.reshape(-1), Z["rt/%d/Rinv" % i]) and np.array_equal(Tinv, Z["rt/%d/Tinv" % i])
