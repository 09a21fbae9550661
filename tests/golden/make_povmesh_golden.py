#!/usr/bin/env python
"""Known answers for the stages after the matcher, from the REFERENCE'S OWN CODE.

oracle/_ref/povmesh_ref (built by oracle/build_ref.sh) is src/wass_stereo/PovMesh.cpp and src/wass_lib/triangulate.hpp of
/root/reference, unmodified, compiled against the header shim in oracle/shim/.  This script feeds it seeded point grids
and point pairs and stores inputs + outputs in tests/golden/povmesh_golden.npz; tests/test_oracle_pipeline.py holds
oracle/pipeline.py (the numpy restatement the GPU kernels are compared with) to these vectors.  Run it here, in the build
container (the GPU box has no /root/reference):

    bash oracle/build_ref.sh && python tests/golden/make_povmesh_golden.py
"""
import os
import struct
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.path.join(ROOT, "oracle", "_ref", "povmesh_ref")


def sea_grid(W, H, seed, holes=0.06, island=True, outliers=12, plane=(0.05, -0.42, 0.9, 8.0), quantum=1.0 / 1024):
    """A tilted sea plane seen from an oblique camera + waves, as a W x H point grid with holes, one detached island
    (a second connected component in z) and a few gross outliers.  Coordinates are multiples of `quantum` (the fixture
    compresses, and text round trips are exact)."""
    rng = np.random.default_rng(seed)
    a, b, c, d = plane
    n = np.array([a, b, c]) / np.linalg.norm([a, b, c])
    v, u = np.mgrid[0:H, 0:W].astype(np.float64)
    X = (u - W / 2) * (30.0 / W)
    Y = (v - H / 2) * (20.0 / H) + 2.0
    Z = (-d - n[0] * X - n[1] * Y) / n[2]
    Z = -Z if Z.mean() < 0 else Z
    Z = Z + 0.15 * np.sin(X * 1.3) * np.cos(Y * 0.9) + rng.normal(0, 0.01, Z.shape)
    valid = rng.random((H, W)) > holes
    if island:
        ys, xs = slice(H // 8, H // 8 + H // 6), slice(W // 10, W // 10 + W // 7)
        Z[ys, xs] += 3.0                                   # far beyond any z-gap percentile: its own component
        valid[ys.start - 1, xs] = valid[ys.stop, xs] = False
    for _ in range(outliers):
        Z[rng.integers(0, H), rng.integers(0, W)] += rng.choice([-1, 1]) * rng.uniform(1.0, 4.0)
    P = np.stack([X, Y, Z], axis=-1)
    P = np.round(P / quantum) * quantum
    grey = rng.integers(20, 250, (H, W)).astype(np.uint8)
    return valid, P, grey


def run_mesh(valid, P, grey, seed, rounds, thr, zpct, maxdist, bounds=(-9999, 9999, -9999, 9999), config=None):
    H, W = valid.shape
    with tempfile.TemporaryDirectory() as td:
        with open(os.path.join(td, "in.bin"), "wb") as f:
            f.write(struct.pack("<ii", W, H))
            f.write(valid.astype(np.uint8).tobytes())
            f.write(np.ascontiguousarray(P, np.float64).tobytes())
            f.write(grey.astype(np.uint8).tobytes())
        cmd = [REF, "mesh", os.path.join(td, "in.bin"), td, str(seed), str(rounds), repr(thr), repr(zpct), repr(maxdist)] + [repr(float(b)) for b in bounds]
        if config:
            with open(os.path.join(td, "cfg.txt"), "w") as f:
                f.write(config)
            cmd.append(os.path.join(td, "cfg.txt"))
        subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL)
        res = {}
        for line in open(os.path.join(td, "result.txt")):
            k, *vals = line.split()
            res[k] = np.array([float(x) for x in vals])
        out = {k: v for k, v in res.items()}
        for m in ("mask_component", "mask_crop1", "mask_final"):
            p = os.path.join(td, m + ".u8")
            if os.path.exists(p):
                out[m] = np.fromfile(p, np.uint8).reshape(H, W)
        for fn in ("mesh_cam.xyzC", "mesh_cam.xyzbin", "mesh.ply"):
            out[fn.replace(".", "_")] = np.fromfile(os.path.join(td, fn), np.uint8)
    return out


def run_tri(items):
    with tempfile.TemporaryDirectory() as td:
        with open(os.path.join(td, "in.bin"), "wb") as f:
            f.write(struct.pack("<i", items.shape[0]))
            f.write(np.ascontiguousarray(items, np.float64).tobytes())
        subprocess.run([REF, "tri", os.path.join(td, "in.bin"), os.path.join(td, "out.bin")], check=True)
        return np.fromfile(os.path.join(td, "out.bin"), np.float64).reshape(-1, 3)


def run_rt(plane):
    txt = subprocess.run([REF, "rt"] + [repr(float(v)) for v in plane], check=True, capture_output=True, text=True).stdout
    out = {}
    for line in txt.splitlines():
        k, *vals = line.split()
        out[k] = np.array([float(x) for x in vals])
    return out


def main():
    if not os.path.exists(REF):
        sys.exit("build oracle/_ref/povmesh_ref first: bash oracle/build_ref.sh")
    z = {}
    cases = [
        # name, W, H, seed, grid kwargs, (srand seed, rounds, ransac thr, zgap pct, plane max dist), bounds, config
        ("sea", 96, 72, 11, {}, (1234, 60, 0.25, 99.0, 0.6), (-9999, 9999, -9999, 9999), None),
        ("sea_bounds_uniform", 80, 60, 12, {"holes": 0.1}, (77, 40, 0.3, 97.5, 0.5), (-8.0, 9.5, -3.0, 9999),
         "PLANE_WEIGHT_PROPORTIONAL_TO_DISTANCE=false\nPLANE_REFINEMENT_MAX_DISTANCE=22.5\n"),
        ("sea_central_third", 64, 48, 13, {"island": False}, (5, 30, 0.3, 99.0, 1.5), (-9999, 9999, -9999, 9999),
         "PLANE_USE_CENTRAL_THIRD_ONLY=true\n"),
        ("noise_ransac_fails", 40, 30, 14, {"holes": 0.3}, (9, 25, 0.01, 90.0, 1.5), (-9999, 9999, -9999, 9999), None),
    ]
    names = []
    for name, W, H, seed, kw, (rs, rounds, thr, zpct, maxd), bounds, cfg in cases:
        valid, P, grey = sea_grid(W, H, seed, **kw)
        if name == "noise_ransac_fails":
            rng = np.random.default_rng(99)
            P = np.round(rng.uniform(-5, 5, P.shape) * 1024) / 1024      # no plane: RANSAC must report failure, softly
        out = run_mesh(valid, P, grey, rs, rounds, thr, zpct, maxd, bounds, cfg)
        z[name + "/valid"] = valid
        z[name + "/p3d"] = P
        z[name + "/grey"] = grey
        z[name + "/args"] = np.array([rs, rounds, thr, zpct, maxd] + list(bounds), np.float64)
        z[name + "/config"] = np.frombuffer((cfg or "").encode(), np.uint8)
        for k, v in out.items():
            z[name + "/" + k] = v
        names.append(name)
        print(name, {k: (v if v.size <= 4 else v.shape) for k, v in out.items() if not k.startswith("mask") and not k.startswith("mesh")})
    # component tie: two components of equal size; the one found first by the reference's column-major rescan wins
    valid = np.zeros((8, 12), bool)
    P = np.zeros((8, 12, 3))
    P[..., 0], P[..., 1] = np.meshgrid(np.arange(12.0), np.arange(8.0))
    valid[5:7, 0:3] = True; P[5:7, 0:3, 2] = 10.0          # found first column-major (u = 0), second row-major
    valid[1:3, 6:9] = True; P[1:3, 6:9, 2] = 20.0
    valid[4, 10] = True; P[4, 10, 2] = 30.0
    # a z-gap percentile needs neighbours: both blocks have internal gaps 0, use a fixed gap via percentile of zeros -> 0;
    # so give the blocks a small internal ramp
    P[5:7, 0:3, 2] += np.arange(3) * 0.01
    P[1:3, 6:9, 2] += np.arange(3) * 0.01
    out = run_mesh(valid, P, np.full(valid.shape, 100, np.uint8), 1, 3, 0.5, 99.0, 1.5)
    z["tie/valid"], z["tie/p3d"] = valid, P
    for k in ("zgap", "mask_component", "n_component"):
        z["tie/" + k] = out[k]
    print("tie", out["zgap"], out["n_component"])
    # triangulate(p, q, R, T): random rigs around the synthetic geometry of SURVEY 8d
    rng = np.random.default_rng(5)
    items = []
    for _ in range(200):
        ang = rng.normal(0, 0.05, 3)
        Rx = np.array([[1, 0, 0], [0, np.cos(ang[0]), -np.sin(ang[0])], [0, np.sin(ang[0]), np.cos(ang[0])]])
        Ry = np.array([[np.cos(ang[1]), 0, np.sin(ang[1])], [0, 1, 0], [-np.sin(ang[1]), 0, np.cos(ang[1])]])
        Rz = np.array([[np.cos(ang[2]), -np.sin(ang[2]), 0], [np.sin(ang[2]), np.cos(ang[2]), 0], [0, 0, 1]])
        R = Rz @ Ry @ Rx
        T = np.array([1.0, 0, 0]) + rng.normal(0, 0.05, 3)
        Xw = np.array([rng.uniform(-10, 10), rng.uniform(-5, 5), rng.uniform(5, 60)])
        p = Xw[:2] / Xw[2] + rng.normal(0, 1e-3, 2)
        Xc = R @ Xw + T
        q = Xc[:2] / Xc[2] + rng.normal(0, 1e-3, 2)
        items.append(np.concatenate([p, q, R.reshape(-1), T]))
    items = np.array(items)
    z["tri/items"] = items
    z["tri/xyz"] = run_tri(items)
    planes = np.array([[0.05, -0.42, 0.9062, 8.0], [0.0, -0.6, 0.8, 3.5], [-0.3, 0.1, 0.9486832980505138, -12.0]])
    z["rt/planes"] = planes
    for i, pl in enumerate(planes):
        for k, v in run_rt(pl).items():
            z["rt/%d/%s" % (i, k)] = v
    z["names"] = np.array(names)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "povmesh_golden.npz"), **z)
    print("wrote tests/golden/povmesh_golden.npz", os.path.getsize(os.path.join(ROOT, "tests", "golden", "povmesh_golden.npz")), "bytes")


if __name__ == "__main__":
    main()
