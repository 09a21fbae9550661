"""Helpers for the workdir interface of wass_stereo: write the inputs the
reference's stage 4 expects (what wass_prepare + wass_autocalibrate leave behind) and read its outputs."""
import os
import struct
import zlib
import numpy as np


def _xml_matrix(name, m):
    m = np.asarray(m, np.float64)
    if m.ndim == 1:
        m = m.reshape(-1, 1)
    data = " ".join(repr(float(v)) for v in m.reshape(-1))
    return ("<?xml version=\"1.0\"?>\n<opencv_storage>\n<%s type_id=\"opencv-matrix\">\n  <rows>%d</rows>\n  <cols>%d</cols>\n"
            "  <dt>d</dt>\n  <data>\n    %s</data></%s>\n</opencv_storage>\n" % (name, m.shape[0], m.shape[1], data, name))


def write_png_gray(path, img):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    raw = b"".join(b"\x00" + img[y].tobytes() for y in range(h))

    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xFFFFFFFF)
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 0, 0, 0, 0)) +
                chunk(b"IDAT", zlib.compress(raw, 3)) + chunk(b"IEND", b""))


def write_workdir(path, cam0_img, cam1_img, K0, K1, R, T):
    """cam0 = undistorted/00000000.png, cam1 = undistorted/00000001.png; X_cam1 = R X_cam0 + T."""
    os.makedirs(os.path.join(path, "undistorted"), exist_ok=True)
    for name, m in (("ext_R", R), ("ext_T", T), ("intrinsics_00000000", K0), ("intrinsics_00000001", K1)):
        with open(os.path.join(path, name + ".xml"), "w") as f:
            f.write(_xml_matrix(name.split("_")[0] if name.startswith("ext") else "intr", m))
    write_png_gray(os.path.join(path, "undistorted", "00000000.png"), cam0_img)
    write_png_gray(os.path.join(path, "undistorted", "00000001.png"), cam1_img)


def write_config(path, **overrides):
    with open(path, "w") as f:
        for k, v in overrides.items():
            if isinstance(v, bool):
                v = "true" if v else "false"
            elif isinstance(v, str):
                v = '"%s"' % v
            f.write("%s=%s\n" % (k, v))


def load_camera_mesh(path):
    """Reader of mesh_cam.xyzC exactly as gridding/wassgridsurface/wass_utils.py:22-35 does it."""
    with open(path, "rb") as f:
        n = np.fromfile(f, np.uint32, 1)[0]
        limits = np.fromfile(f, np.float64, 6)
        Rinv = np.fromfile(f, np.float64, 9).reshape(3, 3)
        Tinv = np.fromfile(f, np.float64, 3).reshape(3, 1)
        data = np.fromfile(f, np.uint16, int(n) * 3).reshape(int(n), 3).astype(np.float64)
    data = data / limits[0:3] + limits[3:6]
    return (Rinv @ data.T + Tinv).T


def load_matrix_txt(path):
    return np.loadtxt(path, ndmin=2)
