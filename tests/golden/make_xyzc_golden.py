"""Golden vectors for the consumer side of mesh_cam.xyzC, produced by the REFERENCE's own reader
(/root/reference/gridding/wassgridsurface/wass_utils.py: load_camera_mesh, align_on_sea_plane) -- run in the build
container only (the reference tree does not exist on the GPU box):

    python tests/golden/make_xyzc_golden.py
"""
import importlib.util
import os
import sys
import tempfile
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import pipeline as op   # noqa: E402

spec = importlib.util.spec_from_file_location("wass_utils", "/root/reference/gridding/wassgridsurface/wass_utils.py")
wu = importlib.util.module_from_spec(spec)
spec.loader.exec_module(wu)

rng = np.random.default_rng(5)
Hm, Wm = 40, 64
valid = rng.random((Hm, Wm)) < 0.8
plane = np.array([0.05, -0.62, 0.78, -3.1]); plane[:3] /= np.linalg.norm(plane[:3])
mean_plane = np.array([0.04, -0.60, 0.80, -3.0]); mean_plane[:3] /= np.linalg.norm(mean_plane[:3])
u, v = np.meshgrid(np.arange(Wm), np.arange(Hm))
x = (u - Wm / 2) * 0.3 + rng.normal(0, 0.01, u.shape)
y = (v - Hm / 2) * 0.3 + rng.normal(0, 0.01, u.shape)
z = -(plane[0] * x + plane[1] * y + plane[3]) / plane[2] + 0.1 * np.sin(x) + rng.normal(0, 0.02, u.shape)
p3d = np.stack([x, y, z], axis=-1)
buf = op.xyz_compressed_bytes(valid, p3d, plane)
with tempfile.NamedTemporaryFile(suffix=".xyzC", delete=False) as f:
    f.write(buf)
    name = f.name
mesh = wu.load_camera_mesh(name)
aligned = wu.align_on_sea_plane(mesh, mean_plane) * 2.5
os.unlink(name)
np.savez_compressed(os.path.join(HERE, "xyzc_golden.npz"), xyzc=np.frombuffer(buf, np.uint8), mean_plane=mean_plane,
                    baseline=2.5, mesh_cam=mesh, aligned=aligned, valid=valid, p3d=p3d, plane=plane)
print("points", mesh.shape, "file bytes", len(buf))
