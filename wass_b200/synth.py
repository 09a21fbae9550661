"""Seeded synthetic rectified stereo pairs (the workload generator of SURVEY.md §8d).

The reference ships no test data (its test set is a download, doc/.../testing.html.md:17-19), so
benchmarks and parity tests use this generator.  `seed = frame index`.
"""
import numpy as np


def _bicubic_upsample(coarse, H, W):
    try:
        import cv2
        return cv2.resize(coarse, (W, H), interpolation=cv2.INTER_CUBIC)
    except ImportError:  # pragma: no cover
        from scipy import ndimage
        zy, zx = H / coarse.shape[0], W / coarse.shape[1]
        return ndimage.zoom(coarse, (zy, zx), order=3)[:H, :W]


def _remap_linear(tex, mapx, mapy):
    try:
        import cv2
        return cv2.remap(tex, mapx.astype(np.float32), mapy.astype(np.float32), cv2.INTER_LINEAR,
                         borderMode=cv2.BORDER_REPLICATE)
    except ImportError:  # pragma: no cover
        from scipy import ndimage
        return ndimage.map_coordinates(tex, [mapy, mapx], order=1, mode="nearest")


def disparity_field(W, H, D, d0=4.0):
    """d(x,y) = d0 + (D-8-d0)*y/H + 3 sin(x/40): far-to-near ramp plus ripples."""
    y, x = np.mgrid[0:H, 0:W].astype(np.float32)
    return d0 + (D - 8.0 - d0) * y / H + 3.0 * np.sin(x / 40.0)


def make_pair(W, H, D, seed=0, d0=4.0, noise=2.0, lo=80, hi=176):
    """Returns (right, left, d_true): `right` is WASS's right image = SGBM img1 (reference view),
    `left` is WASS's left image = SGBM img2, with right(x) ~ left(x - d)  (wass_stereo.cpp:837)."""
    rng = np.random.default_rng(seed)
    Wt = W + 2 * D
    # grey range [lo,hi) keeps max(C)+P2 <= 32767 at WASS defaults, the domain in which
    # cv2.StereoSGBM is reproduced bit-exactly (full-range 0..255 textures leave it)
    coarse = rng.integers(lo, hi, (H // 6 + 2, Wt // 6 + 2)).astype(np.float32)
    tex = np.clip(_bicubic_upsample(coarse, H, Wt), 0, 254).astype(np.float32)
    d = disparity_field(W, H, D, d0)
    y, x = np.mgrid[0:H, 0:W].astype(np.float32)
    # the disparity field is defined on the RIGHT (reference) view: right(x) = left(x - d(x,y))
    left = tex[:, D:D + W]
    right = _remap_linear(tex, x + D - d, y)
    right = np.clip(right + rng.normal(0, noise, right.shape), 0, 254).astype(np.uint8)
    left = np.clip(left + rng.normal(0, noise, left.shape), 0, 254).astype(np.uint8)
    return right, left, d


def pad_for_sgbm(right, left, num_disp, disparity_offset=0):
    """Zero-pad exactly as sgbm_dense_stereo does (wass_stereo.cpp:801-831).
    Returns (img1=padded right, img2=padded left) ready for compute(img1,img2)."""
    H, W = right.shape
    off = max(disparity_offset, 0)
    comp = max(-disparity_offset, 0)
    Wp = W + num_disp + off
    li = np.zeros((H, Wp), np.uint8)
    li[:, num_disp + off - comp: Wp - comp] = left
    ri = np.zeros((H, Wp), np.uint8)
    ri[:, num_disp:num_disp + W] = right
    return ri, li


def make_calibration(W, H):
    """Unit-baseline rectified rig of SURVEY.md §8d: K0=K1, R=I, T=(1,0,0)."""
    f = float(W)
    K = np.array([[f, 0, W / 2.0], [0, f, H / 2.0], [0, 0, 1]], np.float64)
    return dict(K0=K.copy(), K1=K.copy(), R=np.eye(3), T=np.array([1.0, 0.0, 0.0]))
