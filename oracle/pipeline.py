"""CPU restatement of the wass_stereo stages around the matcher -- TEST INFRASTRUCTURE ONLY.

Each function cites the reference lines it follows (paths relative to the reference root).  numpy for the
array arithmetic, explicit Python loops only where the reference's loop order matters (small inputs).
OpenCV calls of the reference are restated in numpy (cv::solve 3x3, cv::SVD) and pinned against cv2
in tests/test_oracle_pipeline.py.  Parity of this file is pinned by fixtures generated with
tests/golden/make_golden.py (cv2 where the reference calls OpenCV) and by analytic ground truth of the
synthetic generator; the reference executable itself cannot be built here (no OpenCV C++/Boost).
"""
import ctypes
import struct
import numpy as np

F32 = np.float32


# ----------------------------------------------------------------------------------------------
# disparity clean-up  (src/wass_stereo/wass_stereo.cpp:617-733, 853-928)
# ----------------------------------------------------------------------------------------------
def clean_and_convert_disparity(disp16, mindisp, num_disp, disp_offset, scale):
    """wass_stereo.cpp:714-733"""
    d = disp16.astype(F32) / F32(16.0)
    keep = ~((d <= F32(mindisp)) | (d > F32(num_disp)))
    v = (d + F32(disp_offset)).astype(F32)
    out = np.zeros(disp16.shape, F32)
    out[keep] = (v.astype(np.float64) * np.float64(scale)).astype(F32)[keep]
    return out


def matrix_dilate_zero(src):
    """wass_stereo.cpp:617-662, including the one-column shift of the output pointer."""
    out = src.copy()
    H, W = src.shape
    if H < 3 or W < 3:
        return out
    i0, i1 = 1, H - 1          # rows 1..H-2
    k = np.arange(0, W - 2)    # written column k, neighbourhood centred on k+1
    T = src[i0 - 1:i1 - 1]
    B = src[i0 + 1:i1 + 1]
    Cc = src[i0:i1]
    # accumulation order of the reference: tm1, tp1, t, bm1, bp1, b, cm1, cp1
    terms = [T[:, k], T[:, k + 2], T[:, k + 1], B[:, k], B[:, k + 2], B[:, k + 1], Cc[:, k], Cc[:, k + 2]]
    avg = np.zeros((i1 - i0, k.size), F32)
    num = np.zeros((i1 - i0, k.size), np.int32)
    for t in terms:
        pos = t > 0
        avg = np.where(pos, (avg + t).astype(F32), avg)
        num += pos
    centre_zero = out[i0:i1][:, k] == 0
    wr = centre_zero & (num > 1)
    res = (avg / np.maximum(num, 1).astype(F32)).astype(F32)
    blk = out[i0:i1, 0:W - 2]
    blk[wr] = res[wr]
    return out


def matrix_erode_zero(src):
    """wass_stereo.cpp:665-711"""
    out = src.copy()
    H, W = src.shape
    if H >= 3 and W >= 3:
        z = src == 0
        zt, zb, zc = z[0:H - 2], z[2:H], z[1:H - 1]
        anyz = (zt[:, 1:W - 1] | zt[:, 0:W - 2] | zt[:, 2:W] | zb[:, 1:W - 1] | zb[:, 0:W - 2] | zb[:, 2:W] |
                zc[:, 0:W - 2] | zc[:, 2:W])
        inner = out[1:H - 1, 1:W - 1]
        inner[anyz] = 0
    if H >= 3:
        out[1:H - 1, 0] = 0
        out[1:H - 1, W - 1] = 0
    out[0, :] = 0
    out[H - 1, :] = 0
    return out


# -- cv::resize as wass_stereo uses it at DENSE_SCALE != 1 (wass_stereo.cpp:790-795, 903-904) -----------------------------
# OpenCV's own resize code (modules/imgproc/src/resize.cpp: resizeGeneric_ with HResizeCubic / VResizeCubic, resizeNN),
# restated from its published algorithm and pinned bit for bit against cv2 4.13 with Intel IPP switched off
# (cv2.setUseOptimized(False); cv2.ipp.setUseIPP(False)) in tests/test_resize.py.  The cv2 wheel's DEFAULT path hands
# INTER_CUBIC to IPP, whose arithmetic differs (u8: +-1 on 2-6 % of pixels); conda-forge's libopencv, which the reference
# pins (meta.yaml), is built without IPP, so the non-IPP code is the reference's own arithmetic.
def _cubic_coeffs(x):
    """interpolateCubic, A = -0.75, float32, in OpenCV's operation order."""
    x = x.astype(F32)
    A, one = F32(-0.75), F32(1)
    xp = (x + one).astype(F32)
    c0 = ((((A * xp).astype(F32) - F32(5) * A).astype(F32) * xp).astype(F32) + F32(8) * A).astype(F32)
    c0 = ((c0 * xp).astype(F32) - F32(4) * A).astype(F32)
    c1 = (((((A + F32(2)) * x).astype(F32) - (A + F32(3))).astype(F32) * x).astype(F32) * x).astype(F32) + one
    xm = (one - x).astype(F32)
    c2 = (((((A + F32(2)) * xm).astype(F32) - (A + F32(3))).astype(F32) * xm).astype(F32) * xm).astype(F32) + one
    c3 = (((one - c0).astype(F32) - c1).astype(F32) - c2).astype(F32)
    return np.stack([c0, c1.astype(F32), c2.astype(F32), c3], -1).astype(F32)


def _cubic_taps(dn, sn, inv_scale):
    """Source indices (replicate-clamped) and float32 coefficients of the 4 taps of every destination coordinate."""
    scale = 1.0 / inv_scale
    f = ((np.arange(dn) + 0.5) * scale - 0.5).astype(F32)          # the coordinate is rounded to float32 first
    s = np.floor(f).astype(np.int64)
    fr = (f - s.astype(F32)).astype(F32)
    idx = np.clip(s[:, None] + np.arange(-1, 3)[None, :], 0, sn - 1)
    return idx, _cubic_coeffs(fr)


def resize_size(n, factor):
    """saturate_cast<int>(n * factor): round half to even."""
    return int(np.rint(n * factor))


def resize_cubic_u8(src, fx, fy):
    """cv::resize(src, dst, Size(), fx, fy, INTER_CUBIC) on CV_8UC1: 11-bit fixed-point coefficients, int32 horizontal pass;
    vertical pass in float32 (S0*b0 + (S1*b1 + (S2*b2 + S3*b3)), round half to even) on the columns the 8-lane baseline
    SIMD loop covers and in 22-bit fixed point ((sum + 2^21) >> 22) on the remaining dw % 8 columns."""
    src = np.asarray(src, np.uint8)
    sh, sw = src.shape
    dw, dh = resize_size(sw, fx), resize_size(sh, fy)
    xi, xc = _cubic_taps(dw, sw, fx)
    yi, yc = _cubic_taps(dh, sh, fy)
    sat = lambda v: np.clip(np.rint(v), -32768, 32767).astype(np.int64)
    ia, ib = sat(xc * F32(2048)), sat(yc * F32(2048))
    S = src.astype(np.int64)
    Hh = np.zeros((sh, dw), np.int64)
    for k in range(4):
        Hh += S[:, xi[:, k]] * ia[None, :, k]
    acc = np.zeros((dh, dw), np.int64)
    for k in range(4):
        acc += Hh[yi[:, k]] * ib[:, k][:, None]
    fixed = np.clip((acc + (1 << 21)) >> 22, 0, 255).astype(np.uint8)
    b = (ib.astype(F32) * F32(1.0 / (2048 * 2048))).astype(F32)
    R = [Hh[yi[:, k]].astype(F32) for k in range(4)]
    t = (R[3] * b[:, 3][:, None]).astype(F32)
    for k in (2, 1, 0):
        t = ((R[k] * b[:, k][:, None]).astype(F32) + t).astype(F32)
    out = np.clip(np.rint(t), 0, 255).astype(np.uint8)
    kx = dw // 8 * 8
    out[:, kx:] = fixed[:, kx:]
    return out


def resize_cubic_f32(src, dw, dh):
    """cv::resize(src, dst, Size(dw,dh), 0, 0, INTER_CUBIC) on CV_32FC1: horizontal ((S0*a0 + S1*a1) + S2*a2) + S3*a3,
    vertical S0*b0 + (S1*b1 + (S2*b2 + S3*b3)) on the 4-lane SIMD columns, left to right on the remaining dw % 4."""
    src = np.asarray(src, F32)
    sh, sw = src.shape
    xi, xc = _cubic_taps(dw, sw, dw / sw)
    yi, yc = _cubic_taps(dh, sh, dh / sh)
    Hh = (src[:, xi[:, 0]] * xc[None, :, 0]).astype(F32)
    for k in (1, 2, 3):
        Hh = (Hh + (src[:, xi[:, k]] * xc[None, :, k]).astype(F32)).astype(F32)
    R = [Hh[yi[:, k]] for k in range(4)]
    t = (R[3] * yc[:, 3][:, None]).astype(F32)
    for k in (2, 1, 0):
        t = ((R[k] * yc[:, k][:, None]).astype(F32) + t).astype(F32)
    u = (R[0] * yc[:, 0][:, None]).astype(F32)
    for k in (1, 2, 3):
        u = (u + (R[k] * yc[:, k][:, None]).astype(F32)).astype(F32)
    kx = dw // 4 * 4
    t[:, kx:] = u[:, kx:]
    return t


def resize_nearest(src, dw, dh):
    """cv::resize(..., INTER_NEAREST): sx = min(floor(x / (dw/sw)), sw-1) in double."""
    sh, sw = src.shape
    ifx, ify = 1.0 / (dw / sw), 1.0 / (dh / sh)
    xs = np.minimum(np.floor(np.arange(dw) * ifx).astype(np.int64), sw - 1)
    ys = np.minimum(np.floor(np.arange(dh) * ify).astype(np.int64), sh - 1)
    return src[ys][:, xs]


def clahe_u8(src, clip_limit, tiles):
    """cv::createCLAHE(clip_limit, Size(tiles,tiles))->apply(src) on CV_8UC1 (src/wass_prepare/wass_prepare.cpp:257-262,
    458-462): OpenCV's published algorithm (modules/imgproc/src/clahe.cpp), pinned bit-exact against cv2 in
    tests/test_prepare.py."""
    src = np.asarray(src, np.uint8)
    H, W = src.shape
    if W % tiles == 0 and H % tiles == 0:
        ext = src
    else:   # both paddings are applied, a whole extra `tiles` where the size already divides
        ext = np.pad(src, ((0, tiles - H % tiles), (0, tiles - W % tiles)), mode="reflect")
    tw, th = ext.shape[1] // tiles, ext.shape[0] // tiles
    area = tw * th
    lut_scale = F32(255) / F32(area)
    clip = max(int(clip_limit * area / 256), 1) if clip_limit > 0 else 0
    luts = np.zeros((tiles, tiles, 256), np.uint8)
    for j in range(tiles):
        for i in range(tiles):
            h = np.bincount(ext[j * th:(j + 1) * th, i * tw:(i + 1) * tw].ravel(), minlength=256).astype(np.int64)
            if clip > 0:
                clipped = int(np.maximum(h - clip, 0).sum())
                h = np.minimum(h, clip)
                batch = clipped // 256
                residual = clipped - batch * 256
                h += batch
                if residual:
                    step = max(256 // residual, 1)
                    k = np.arange(0, 256, step)[:residual]
                    h[k] += 1
            luts[j, i] = np.clip(np.rint((np.cumsum(h).astype(F32) * lut_scale).astype(F32)), 0, 255).astype(np.uint8)

    def blend_coords(n, inv, ntiles):
        f = (np.arange(n).astype(F32) * inv - F32(0.5)).astype(F32)
        t1 = np.floor(f).astype(np.int64)
        a = (f - t1.astype(F32)).astype(F32)
        return np.maximum(t1, 0), np.minimum(t1 + 1, ntiles - 1), a, (F32(1) - a).astype(F32)

    tx1, tx2, xa, xa1 = blend_coords(W, F32(1) / F32(tw), tiles)
    ty1, ty2, ya, ya1 = blend_coords(H, F32(1) / F32(th), tiles)
    v = src.astype(np.int64)
    L = luts.astype(F32)
    top = ((L[ty1[:, None], tx1[None, :], v] * xa1[None, :]).astype(F32) + (L[ty1[:, None], tx2[None, :], v] * xa[None, :]).astype(F32)).astype(F32)
    bot = ((L[ty2[:, None], tx1[None, :], v] * xa1[None, :]).astype(F32) + (L[ty2[:, None], tx2[None, :], v] * xa[None, :]).astype(F32)).astype(F32)
    res = ((top * ya1[:, None]).astype(F32) + (bot * ya[:, None]).astype(F32)).astype(F32)
    return np.clip(np.rint(res), 0, 255).astype(np.uint8)


# -- custom rectifier (src/wass_stereo/stereorectify.cpp:57-244, USE_CUSTOM_STEREORECTIFY) ------------------------------
def _custom_rect_functional(K0, K1, R, T):
    """HFunctional (stereorectify.cpp:70-137): returns f(angle_deg) -> (max(v1, v2), H0, H1)."""
    K0i, K1i = np.linalg.inv(np.asarray(K0, float)), np.linalg.inv(np.asarray(K1, float))
    Ri = np.asarray(R, float)
    Rv = np.asarray(T, float).ravel() / np.linalg.norm(T)
    N = np.cross(Rv, [0.0, 1.0, 0.0])
    N /= np.linalg.norm(N)
    Rplane = np.stack([Rv, np.cross(Rv, N), N])

    def f(x):
        a = x / 180 * 3.14                       # sic: 3.14
        c, s_ = np.cos(a), np.sin(a)
        Radd = np.array([[1.0, 0, 0], [0, c, -s_], [0, s_, c]])       # Rodrigues of (a, 0, 0)
        H0 = Radd @ Rplane @ K0i
        H1 = Radd @ Rplane @ Ri @ K1i
        H0 = H0 / H0[2, 2]
        H1 = H1 / H1[2, 2]
        v = max(H0[2, 0] ** 2 + H0[2, 1] ** 2, H1[2, 0] ** 2 + H1[2, 1] ** 2)
        return v, H0 / np.cbrt(np.linalg.det(H0)), H1 / np.cbrt(np.linalg.det(H1))
    return f


def custom_rectify_best_angle(K0, K1, R, T, lo=-60.0, hi=60.0):
    """The minimiser of the functional near 0 by an independent method (dense scan + golden section): what
    cv::DownhillSolver converges to from (0,0) with step -0.5 (cv2 does not export that solver: parity unpinned)."""
    f = _custom_rect_functional(K0, K1, R, T)
    xs = np.linspace(lo, hi, 2401)
    vals = np.array([f(x)[0] for x in xs])
    # the local minimum a downhill walk from 0 reaches
    i = int(np.argmin(np.abs(xs)))
    while 0 < i < len(xs) - 1 and (vals[i - 1] < vals[i] or vals[i + 1] < vals[i]):
        i = i - 1 if vals[i - 1] < vals[i + 1] else i + 1
    a, b = xs[max(i - 1, 0)], xs[min(i + 1, len(xs) - 1)]
    g = (np.sqrt(5) - 1) / 2
    for _ in range(200):
        c, d = b - g * (b - a), a + g * (b - a)
        if f(c)[0] < f(d)[0]:
            b = d
        else:
            a = c
    return 0.5 * (a + b)


def stereo_rectify_custom(K0, K1, R, T, width, height, angle):
    """stereoRectifyUndistorted for a GIVEN baseline angle (degrees): H0, H1 and the ROI (x, y, w, h)."""
    f = _custom_rect_functional(K0, K1, R, T)
    _, H0, H1 = f(angle)
    pts = np.array([[0, width, width, 0], [0, 0, height, height], [1, 1, 1, 1]], float)

    def corners(H):
        q = H @ pts
        return q[0] / q[2], q[1] / q[2]

    def rect(cx, cy):
        ax, ay, bx, by = min(cx[0], cx[3]), min(cy[0], cy[1]), max(cx[1], cx[2]), max(cy[2], cy[3])
        return min(ax, bx), min(ay, by), abs(bx - ax), abs(by - ay)
    r0, r1 = rect(*corners(H0)), rect(*corners(H1))
    top, bottom = min(r0[1], r1[1]), max(r0[1] + r0[3], r1[1] + r1[3])
    out = []
    for H, r in ((H0, r0), (H1, r1)):
        Tr = np.array([[1, 0, -r[0]], [0, 1, -top], [0, 0, 1.0]])
        Sc = np.diag([width / r[2], height / (bottom - top), 1.0])
        Hn = Sc @ Tr @ H
        out.append(Hn / np.cbrt(np.linalg.det(Hn)))
    (x0, y0), (x1, y1) = corners(out[0]), corners(out[1])
    xs, ys = np.sort(np.concatenate([x0, x1])), np.sort(np.concatenate([y0, y1]))
    rx, ry = int(xs[3]), int(ys[3])
    return out[0], out[1], (rx, ry, int(xs[4] - rx), int(ys[4] - ry))


def warp_perspective_u8(src, H):
    """cv::warpPerspective(src, dst, H, src.size()) (INTER_LINEAR, constant 0 border; wass_stereo.cpp:515-516): OpenCV's
    published algorithm -- M = H^-1, coordinates in double per 128-column block (origin term + M*x1), scaled by 32/W,
    rounded half to even, 1/32-pixel bilinear weights in 15-bit fixed point -- bit-exact vs cv2 in tests/test_custom_rectify.py."""
    src = np.asarray(src, np.uint8)
    sh, sw = src.shape
    M = np.linalg.inv(np.asarray(H, float))
    S = np.asarray(H, float).ravel()        # cv::invert's closed form for 3x3
    d = S[0] * (S[4] * S[8] - S[5] * S[7]) - S[1] * (S[3] * S[8] - S[5] * S[6]) + S[2] * (S[3] * S[7] - S[4] * S[6])
    idet = 1.0 / d
    M = np.array([(S[4] * S[8] - S[5] * S[7]) * idet, (S[2] * S[7] - S[1] * S[8]) * idet, (S[1] * S[5] - S[2] * S[4]) * idet,
                  (S[5] * S[6] - S[3] * S[8]) * idet, (S[0] * S[8] - S[2] * S[6]) * idet, (S[2] * S[3] - S[0] * S[5]) * idet,
                  (S[3] * S[7] - S[4] * S[6]) * idet, (S[1] * S[6] - S[0] * S[7]) * idet, (S[0] * S[4] - S[1] * S[3]) * idet])
    bh0 = min(32, sh)
    bw0 = min(64 * 64 // bh0, sw)
    xs, ys = np.arange(sw), np.arange(sh)
    xb = (xs // bw0) * bw0
    x1 = xs - xb
    X0 = (M[0] * xb[None, :] + M[1] * ys[:, None]) + M[2]
    Y0 = (M[3] * xb[None, :] + M[4] * ys[:, None]) + M[5]
    W0 = (M[6] * xb[None, :] + M[7] * ys[:, None]) + M[8]
    Wv = W0 + M[6] * x1[None, :]
    with np.errstate(divide="ignore", invalid="ignore"):
        Wi = np.where(Wv != 0, 32.0 / Wv, 0.0)
        fX = np.clip((X0 + M[0] * x1[None, :]) * Wi, -2147483648.0, 2147483647.0)
        fY = np.clip((Y0 + M[3] * x1[None, :]) * Wi, -2147483648.0, 2147483647.0)
    X, Y = np.rint(fX).astype(np.int64), np.rint(fY).astype(np.int64)
    ix, iy = np.clip(X >> 5, -32768, 32767), np.clip(Y >> 5, -32768, 32767)
    fx, fy = X & 31, Y & 31
    w = np.stack([(32 - fx) * (32 - fy) * 32, fx * (32 - fy) * 32, (32 - fx) * fy * 32, fx * fy * 32], -1)
    w[..., 0] = np.minimum(w[..., 0], 32767)

    def px(y, x):
        ok = (x >= 0) & (x < sw) & (y >= 0) & (y < sh)
        return np.where(ok, src[np.clip(y, 0, sh - 1), np.clip(x, 0, sw - 1)], 0).astype(np.int64)
    v = px(iy, ix) * w[..., 0] + px(iy, ix + 1) * w[..., 1] + px(iy + 1, ix) * w[..., 2] + px(iy + 1, ix + 1) * w[..., 3]
    return np.clip((v + (1 << 14)) >> 15, 0, 255).astype(np.uint8)


def dense_input_resize(crop, dense_scale):
    """wass_stereo.cpp:788-797: x only when enlarging, both axes when shrinking."""
    if dense_scale > 1.0:
        return resize_cubic_u8(crop, dense_scale, 1.0)
    if dense_scale < 1.0:
        return resize_cubic_u8(crop, dense_scale, dense_scale)
    return crop


def postprocess_disparity(disp16_roi, mindisp, num_disp, disparity_offset=0, dense_scale=1.0,
                          dilate_steps=1, erosion_steps=2, out_size=None):
    """wass_stereo.cpp:853-928.  disp16_roi has the size of the (resized) matcher input; out_size = (rows, cols) of
    roi_comb_right, needed when DENSE_SCALE != 1 (at 1 both cv::resize calls are exact copies)."""
    off = max(disparity_offset, 0)
    d = clean_and_convert_disparity(disp16_roi, mindisp, num_disp, off, 1.0 / dense_scale)
    for _ in range(max(dilate_steps, 0)):
        d = matrix_dilate_zero(d)
    for _ in range(max(erosion_steps, 0)):
        d = matrix_erode_zero(d)
    if dense_scale == 1.0 and (out_size is None or tuple(out_size) == d.shape):
        nn, cub = d, d
    else:
        nn = resize_nearest(d, out_size[1], out_size[0])
        cub = resize_cubic_f32(d, out_size[1], out_size[0])
    nn = matrix_erode_zero(nn)
    out = cub.copy()
    out[nn == 0] = 0
    return out


# ----------------------------------------------------------------------------------------------
# triangulation  (wass_stereo.cpp:299-324, 1039-1386; src/wass_lib/triangulate.hpp:26-72)
# ----------------------------------------------------------------------------------------------
def median_f32(img, ksize):
    """cv::medianBlur on float32 (ksize 3 or 5, replicate border), wass_stereo.cpp:941-945."""
    r = ksize // 2
    p = np.pad(np.asarray(img, np.float32), r, mode="edge")
    H, W = img.shape
    st = np.stack([p[i:i + H, j:j + W] for i in range(ksize) for j in range(ksize)], 0)
    return np.sort(st, 0)[ksize * ksize // 2]


def gradient_mask(img, threshold):
    """wass_stereo.cpp:947-963: zero where Sobel gx^2 + gy^2 > threshold.  Float operation order of cv::Sobel (3x3,
    BORDER_REFLECT_101), found by matching cv2 bit for bit: gx = ((d[y-1]+d[y+1]) + 2 d[y]), d = p[x+1]-p[x-1];
    gy = s[y+1]-s[y-1], s = ((p[x-1]+p[x+1]) + 2 p[x])."""
    f = np.float32
    a = np.asarray(img, f)
    p = np.pad(a, 1, mode="reflect")
    d = (p[:, 2:] - p[:, :-2]).astype(f)
    gx = ((d[:-2] + d[2:]).astype(f) + (f(2) * d[1:-1]).astype(f)).astype(f)
    sm = ((p[:, :-2] + p[:, 2:]).astype(f) + (f(2) * p[:, 1:-1]).astype(f)).astype(f)
    gy = (sm[2:] - sm[:-2]).astype(f)
    g2 = ((gx * gx).astype(f) + (gy * gy).astype(f)).astype(f)
    out = a.copy()
    out[g2 > f(threshold)] = 0
    return out


def keep_biggest_component8(img):
    """wass_stereo.cpp:965-985: biggest 8-connected component of the non-zero pixels (cv::connectedComponentsWithStats,
    strict '>' over increasing labels).  cv2 numbers components by their first 2x2 block in block-raster order, which
    decides ties."""
    a = np.asarray(img, np.float32)
    H, W = a.shape
    lab = -np.ones((H, W), np.int64)
    nxt = 0
    for y in range(H):
        for x in range(W):
            if a[y, x] == 0 or lab[y, x] >= 0:
                continue
            stack = [(y, x)]
            lab[y, x] = nxt
            while stack:
                cy, cx = stack.pop()
                for dy in (-1, 0, 1):
                    for dx in (-1, 0, 1):
                        yy, xx = cy + dy, cx + dx
                        if 0 <= yy < H and 0 <= xx < W and a[yy, xx] != 0 and lab[yy, xx] < 0:
                            lab[yy, xx] = nxt
                            stack.append((yy, xx))
            nxt += 1
    out = a.copy()
    if nxt == 0:
        return out
    ys, xs = np.nonzero(lab >= 0)
    area = np.bincount(lab[ys, xs], minlength=nxt)
    key = np.full(nxt, 1 << 62)
    np.minimum.at(key, lab[ys, xs], (ys // 2) * ((W + 1) // 2) + xs // 2)
    best = min(range(nxt), key=lambda i: (-area[i], key[i]))
    out[lab != best] = 0
    return out


def refine_disparity(disp, median_wsize=0, bc_threshold=0):
    """wass_stereo.cpp:941-986 in order."""
    d = np.asarray(disp, np.float32)
    if median_wsize >= 3:
        d = median_f32(d, median_wsize)
    if bc_threshold > 0:
        d = keep_biggest_component8(gradient_mask(d, bc_threshold))
    return d


def solve3_lu(A, b):
    """cv::solve(A,b,x,DECOMP_LU) for 3x3: OpenCV's closed-form (Cramer) fast path, in double."""
    det = (A[0, 0] * (A[1, 1] * A[2, 2] - A[1, 2] * A[2, 1]) - A[0, 1] * (A[1, 0] * A[2, 2] - A[1, 2] * A[2, 0]) +
           A[0, 2] * (A[1, 0] * A[2, 1] - A[1, 1] * A[2, 0]))
    if det == 0.0:
        return np.zeros(3)
    d = 1.0 / det
    t0 = d * (b[0] * (A[1, 1] * A[2, 2] - A[1, 2] * A[2, 1]) - A[0, 1] * (b[1] * A[2, 2] - A[1, 2] * b[2]) +
              A[0, 2] * (b[1] * A[2, 1] - A[1, 1] * b[2]))
    t1 = d * (A[0, 0] * (b[1] * A[2, 2] - A[1, 2] * b[2]) - b[0] * (A[1, 0] * A[2, 2] - A[1, 2] * A[2, 0]) +
              A[0, 2] * (A[1, 0] * b[2] - b[1] * A[2, 0]))
    t2 = d * (A[0, 0] * (A[1, 1] * b[2] - b[1] * A[2, 1]) - A[0, 1] * (A[1, 0] * b[2] - b[1] * A[2, 0]) +
              b[0] * (A[1, 0] * A[2, 1] - A[1, 1] * A[2, 0]))
    return np.array([t0, t1, t2])


def triangulate_point(p, q, R, T):
    """src/wass_lib/triangulate.hpp:26-72 (normal equations of the 4x3 system, same summation order)."""
    Af = np.array([-1.0, 0.0, p[0],
                   0.0, -1.0, p[1],
                   q[0] * R[2, 0] - R[0, 0], q[0] * R[2, 1] - R[0, 1], q[0] * R[2, 2] - R[0, 2],
                   q[1] * R[2, 0] - R[1, 0], q[1] * R[2, 1] - R[1, 1], q[1] * R[2, 2] - R[1, 2]])
    Bf = np.array([0.0, 0.0, T[0] - T[2] * q[0], T[1] - T[2] * q[1]])
    A = np.empty(9)
    A[0] = Af[0] * Af[0] + Af[3] * Af[3] + Af[6] * Af[6] + Af[9] * Af[9]
    A[1] = Af[0] * Af[1] + Af[3] * Af[4] + Af[10] * Af[9] + Af[6] * Af[7]
    A[2] = Af[0] * Af[2] + Af[3] * Af[5] + Af[11] * Af[9] + Af[6] * Af[8]
    A[3] = A[1]
    A[4] = Af[1] * Af[1] + Af[10] * Af[10] + Af[4] * Af[4] + Af[7] * Af[7]
    A[5] = Af[10] * Af[11] + Af[1] * Af[2] + Af[4] * Af[5] + Af[7] * Af[8]
    A[6] = A[2]
    A[7] = A[5]
    A[8] = Af[11] * Af[11] + Af[2] * Af[2] + Af[5] * Af[5] + Af[8] * Af[8]
    b = np.empty(3)
    b[0] = Af[0] * Bf[0] + Af[3] * Bf[1] + Af[6] * Bf[2] + Af[9] * Bf[3]
    b[1] = Af[1] * Bf[0] + Af[10] * Bf[3] + Af[4] * Bf[1] + Af[7] * Bf[2]
    b[2] = Af[2] * Bf[0] + Af[11] * Bf[3] + Af[5] * Bf[1] + Af[8] * Bf[2]
    return solve3_lu(A.reshape(3, 3), b)


def unrectify(uv, Kold, Rrect, Pnew):
    """wass_stereo.cpp:299-324 (cv::stereoRectify branch)."""
    x = (uv[0] - Pnew[0, 2]) / Pnew[0, 0]
    y = (uv[1] - Pnew[1, 2]) / Pnew[1, 1]
    Rt = Rrect.T
    v0 = Rt[0, 0] * x + Rt[0, 1] * y + Rt[0, 2] * 1.0
    v1 = Rt[1, 0] * x + Rt[1, 1] * y + Rt[1, 2] * 1.0
    v2 = Rt[2, 0] * x + Rt[2, 1] * y + Rt[2, 2] * 1.0
    v0 /= v2
    v1 /= v2
    return np.array([v0 * Kold[0, 0] + Kold[0, 2], v1 * Kold[1, 1] + Kold[1, 2]])


def unrectify_homography(uv, Hinv):
    """wass_stereo.cpp:301-305: (HLi | HRi) * (u, v, 1), cv::Matx product (row sums from the left), then the division."""
    r = [(Hinv[i, 0] * uv[0] + Hinv[i, 1] * uv[1]) + Hinv[i, 2] * 1.0 for i in range(3)]
    return np.array([r[0] / r[2], r[1] / r[2]])


def triangulate(disparity, calib, left, right, left_mask=None, right_mask=None, min_angle=20.0,
                bbox=None, discard_burned=True, disparity_compensation=0, dense_scale=1.0, cam_distance=1.0):
    """wass_stereo.cpp:1039-1386.  disparity: float32, full rectified size.  calib: dict with
    K0,K1 (left/right intrinsics after any swap), R,T, R1,R2,P1,P2 (cv::stereoRectify outputs) and
    roi_left, roi_right = (x,y,w,h).  left/right: original (unrectified) uint8 images.
    Returns dict(valid bool [h][w], p3d float64 [h][w][3], color uint8 [h][w], n)."""
    K0, K1, R, T = calib["K0"], calib["K1"], calib["R"], np.asarray(calib["T"], np.float64).reshape(3)
    R1, R2, P1, P2 = calib["R1"], calib["R2"], calib["P1"], calib["P2"]
    rlx, rly, rlw, rlh = calib["roi_left"]
    rrx, rry, rrw, rrh = calib["roi_right"]
    Hl, Wl = left.shape
    Hr, Wr = right.shape
    rect_cols = disparity.shape[1]
    lm = np.ones(left.shape, np.uint8) if left_mask is None else (left_mask > 0).astype(np.uint8)
    rm = np.ones(right.shape, np.uint8) if right_mask is None else (right_mask > 0).astype(np.uint8)
    if discard_burned:
        lm = lm * (1 - (left > 254).astype(np.uint8))
        rm = rm * (1 - (right > 254).astype(np.uint8))
    if bbox is None:
        tl, br = (0.0, 0.0), (float(Wl), float(Hl))
    else:
        tl, br = (bbox[0], bbox[1]), (bbox[2], bbox[3])
    valid = np.zeros((rrh, rrw), bool)
    p3d = np.zeros((rrh, rrw, 3), np.float64)
    color = np.zeros((rrh, rrw), np.uint8)
    n = 0
    for yr in range(rry, rry + rrh):
        for xr in range(rrx, rrx + rrw):
            dv = disparity[yr, xr]
            if not dv > 1:
                continue
            xl = F32(F32(xr - rrx + rlx) - dv)
            xl = F32(np.float64(xl) + disparity_compensation / dense_scale)
            yl = F32(yr)
            if xl < 0 or xl >= rect_cols:
                continue
            if "HLi" in calib:      # USE_CUSTOM_STEREORECTIFY: through the inverse homographies (wass_stereo.cpp:301-305)
                pi = unrectify_homography((np.float64(xl), np.float64(yl)), calib["HLi"])
                qi = unrectify_homography((np.float64(xr), np.float64(yr)), calib["HRi"])
            else:
                pi = unrectify((np.float64(xl), np.float64(yl)), K0, R1, P1)
                qi = unrectify((np.float64(xr), np.float64(yr)), K1, R2, P2)
            skip = False
            if (pi[0] < 1 or pi[0] >= Wl - 1 or pi[1] < 1 or pi[1] >= Hl - 1 or
                    qi[0] < 1 or qi[0] >= Wr - 1 or qi[1] < 1 or qi[1] >= Hr - 1):
                continue  # reference sets skip and later indexes masks; out-of-range there is UB -> we stop here
            p = np.array([(pi[0] - K0[0, 2]) / K0[0, 0], (pi[1] - K0[1, 2]) / K0[1, 1]])
            q = np.array([(qi[0] - K1[0, 2]) / K1[0, 0], (qi[1] - K1[1, 2]) / K1[1, 1]])
            if pi[0] <= tl[0] or pi[1] <= tl[1] or pi[0] >= br[0] or pi[1] >= br[1]:
                skip = True
            if lm[int(pi[1]), int(pi[0])] == 0:
                skip = True
            if rm[int(qi[1]), int(qi[0])] == 0:
                skip = True
            if min_angle > 0:
                a = np.array([p[0], p[1], 1.0])
                n1 = np.sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2])
                d1 = a * (1.0 / n1 if n1 else 0.0)
                b0 = R[0, 0] * q[0] + R[0, 1] * q[1] + R[0, 2] * 1.0 + T[0]
                b1 = R[1, 0] * q[0] + R[1, 1] * q[1] + R[1, 2] * 1.0 + T[1]
                b2 = R[2, 0] * q[0] + R[2, 1] * q[1] + R[2, 2] * 1.0 + T[2]
                n2 = np.sqrt(b0 * b0 + b1 * b1 + b2 * b2)
                s = 1.0 / n2 if n2 else 0.0
                d2 = np.array([b0 * s, b1 * s, b2 * s])
                dot = d1[0] * d2[0] + d1[1] * d2[1] + d1[2] * d2[2]
                ang = abs(np.arccos(min(max(dot, -1.0), 1.0)) * 57.29577951) if abs(dot) <= 1 else float("nan")
                if ang < min_angle:
                    skip = True
            if skip:
                continue
            X = triangulate_point(p, q, R, T)
            dist = np.sqrt(X[0] * X[0] + X[1] * X[1] + X[2] * X[2])
            if dist < cam_distance / 10.0 or X[2] < 1.0:
                continue
            if dist > cam_distance * 200.0 or X[2] > 1e30:
                continue
            u, v = xr - rrx, yr - rry
            valid[v, u] = True
            p3d[v, u] = X
            color[v, u] = right[int(qi[1]), int(qi[0])]
            n += 1
    return dict(valid=valid, p3d=p3d, color=color, n=n)


# ----------------------------------------------------------------------------------------------
# PovMesh  (src/wass_stereo/PovMesh.cpp)
# ----------------------------------------------------------------------------------------------
def zgap_percentile(valid, z, percentile=99.0):
    """PovMesh.cpp:888-926"""
    H, W = valid.shape
    c = valid[1:H, 1:W - 1]
    z0 = z[1:H, 1:W - 1]
    gaps = []
    for dx in (-1, 0, 1):
        nb = valid[0:H - 1, 1 + dx:W - 1 + dx]
        zn = z[0:H - 1, 1 + dx:W - 1 + dx]
        m = c & nb
        gaps.append(np.abs(z0[m] - zn[m]))
    g = np.sort(np.concatenate(gaps))
    if g.size == 0:
        return float("nan")
    idx = int(np.floor(percentile / 100.0 * g.size))
    return float(g[min(idx, g.size - 1)]) if idx < g.size else float("nan")


def biggest_component(valid, z, zgap):
    """PovMesh.cpp:929-987 (+147-203): 4-connected, edge iff |dz| < zgap; biggest wins, ties go to the
    component found first by the column-major rescan.  Returns the new valid mask."""
    from scipy import ndimage  # labels only; equivalence classes via union-find below
    H, W = valid.shape
    parent = np.arange(H * W)

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a

    idx = np.arange(H * W).reshape(H, W)
    eh = valid[:, :-1] & valid[:, 1:] & (np.abs(z[:, :-1] - z[:, 1:]) < zgap)
    ev = valid[:-1, :] & valid[1:, :] & (np.abs(z[:-1, :] - z[1:, :]) < zgap)
    for a, b in list(zip(idx[:, :-1][eh], idx[:, 1:][eh])) + list(zip(idx[:-1, :][ev], idx[1:, :][ev])):
        ra, rb = find(a), find(b)
        if ra != rb:
            parent[max(ra, rb)] = min(ra, rb)
    roots = np.array([find(a) if valid.flat[a] else -1 for a in range(H * W)])
    if not valid.any():
        return valid.copy()
    labs, counts = np.unique(roots[roots >= 0], return_counts=True)
    best = counts.max()
    cand = labs[counts == best]
    # first found in column-major scan = smallest (u*H + v) over the members of each candidate
    vv, uu = np.divmod(np.arange(H * W), W)
    colmajor = uu * H + vv
    firsts = [colmajor[roots == c].min() for c in cand]
    win = cand[int(np.argmin(firsts))]
    return (roots == win).reshape(H, W)


class LibcRand:
    """glibc srand/rand, as the reference uses for RANSAC (wass_stereo.cpp:1864-1872, PovMesh.cpp:680-682)."""

    def __init__(self, seed):
        self.libc = ctypes.CDLL("libc.so.6")
        self.libc.srand(ctypes.c_uint(seed))

    def rand(self):
        return self.libc.rand()


def ransac_draw_triples(rng, W, H, rounds):
    """The draw loop of PovMesh.cpp:678-692.  `cv::Vec2i p(rand()%iW, rand()%iH)` leaves the order of the two calls to
    the compiler; GCC evaluates constructor arguments right to left, so a Linux build of the reference
    draws v BEFORE u (points 1, 2, 3 in statement order).  Pinned: tests/golden/povmesh_golden.npz holds planes from the
    reference's own PovMesh.cpp built with GCC and a fixed seed, and only this order reproduces them.
    Returns int32 [rounds][6] pixel coordinates (u1,v1,u2,v2,u3,v3) of the rounds that pass the
    minimum-distance test (each consumes one round)."""
    out = np.zeros((rounds, 6), np.int32)
    mind = H * 0.01
    r = 0
    while r < rounds:
        c = [0] * 6
        for k in range(3):
            c[2 * k + 1] = rng.rand() % H
            c[2 * k] = rng.rand() % W
        d12 = np.hypot(c[0] - c[2], c[1] - c[3])
        d23 = np.hypot(c[2] - c[4], c[3] - c[5])
        d13 = np.hypot(c[0] - c[4], c[1] - c[5])
        if d12 < mind or d23 < mind or d13 < mind:
            continue
        out[r] = c
        r += 1
    return out


def ransac_find_plane(valid, p3d, triples, threshold):
    """PovMesh.cpp:665-777 given the drawn triples.  Returns (ok, plane[4], best_inliers)."""
    H, W = valid.shape
    P = p3d[valid]
    best, best_n, best_d = 0, np.zeros(3), 0.0
    for t in triples:
        u1, v1, u2, v2, u3, v3 = [int(a) for a in t]
        if not (valid[v1, u1] and valid[v2, u2] and valid[v3, u3]):
            continue
        p1, p2, p3 = p3d[v1, u1], p3d[v2, u2], p3d[v3, u3]
        n = np.cross(p2 - p1, p3 - p1)
        n = n / np.sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2])
        if n[2] < 0:
            n = n * -1.0
        d = -(n[0] * p1[0] + n[1] * p1[1] + n[2] * p1[2])
        dist = np.abs(P[:, 0] * n[0] + P[:, 1] * n[1] + P[:, 2] * n[2] + d)
        k = int((dist < threshold).sum())
        if k > best:
            best, best_n, best_d = k, n, d
    ok = not (best < (H * W) // 10)
    return ok, np.array([best_n[0], best_n[1], best_n[2], best_d]), best


def crop_plane(valid, p3d, plane, threshold):
    """PovMesh.cpp:780-815"""
    d = np.abs(p3d[..., 0] * plane[0] + p3d[..., 1] * plane[1] + p3d[..., 2] * plane[2] + plane[3])
    return valid & (d < threshold)


def refine_plane(valid, p3d, xmin=-9999.0, xmax=9999.0, ymin=-9999.0, ymax=9999.0, max_distance=70.0,
                 weight_by_distance=True, central_third=False):
    """PovMesh.cpp:581-660.  Returns (plane[4], n_inliers)."""
    H, W = valid.shape
    umin, umax = (W // 4, W * 3 // 4) if central_third else (0, W - 1)
    vmin, vmax = (H // 4, H * 2 // 3) if central_third else (0, H - 1)
    sel = np.zeros_like(valid)
    sel[vmin:vmax + 1, umin:umax + 1] = True
    P = p3d
    dist = np.sqrt(P[..., 0] * P[..., 0] + P[..., 1] * P[..., 1] + P[..., 2] * P[..., 2])
    m = valid & sel & (P[..., 0] > xmin) & (P[..., 0] < xmax) & (P[..., 1] > ymin) & (P[..., 1] < ymax) & (dist < max_distance)
    pts = P[m]
    w = dist[m] if weight_by_distance else np.ones(pts.shape[0])
    wsum = w.sum()
    c = (pts * w[:, None]).sum(axis=0) / wsum
    q = pts - c
    A = (q[:, :, None] * q[:, None, :] * w[:, None, None]).sum(axis=0)
    _, _, vt = np.linalg.svd(A)
    n = vt[2]
    n = n / np.sqrt((n * n).sum())
    if n[2] < 0:
        n = -n
    d = -(n[0] * c[0] + n[1] * c[1] + n[2] * c[2])
    return np.array([n[0], n[1], n[2], d]), int(m.sum())


def refinement_inlier_samples(valid, p3d, every=10, xmin=-9999.0, xmax=9999.0, ymin=-9999.0, ymax=9999.0, max_distance=70.0,
                              central_third=False):
    """The points main() writes to plane_refinement_inliers.xyz (wass_stereo.cpp:2077-2085): every `every`-th point passing
    refine_plane's inlier test (PovMesh.cpp:596-618), in grid scan order.  Returns (points[n][3], n_inliers)."""
    H, W = valid.shape
    umin, umax = (W // 4, W * 3 // 4) if central_third else (0, W - 1)
    vmin, vmax = (H // 4, H * 2 // 3) if central_third else (0, H - 1)
    sel = np.zeros_like(valid)
    sel[vmin:vmax + 1, umin:umax + 1] = True
    P = p3d
    dist = np.sqrt(P[..., 0] * P[..., 0] + P[..., 1] * P[..., 1] + P[..., 2] * P[..., 2])
    m = valid & sel & (P[..., 0] > xmin) & (P[..., 0] < xmax) & (P[..., 1] > ymin) & (P[..., 1] < ymax) & (dist < max_distance)
    pts = P[m]                       # boolean indexing walks the grid in scan order
    return pts[::every].copy(), int(m.sum())


def rt_from_plane(a, b, c, d):
    """PovMesh.cpp:1044-1074"""
    q = (1 - c) / (a * a + b * b)
    R = np.array([[1 - a * a * q, -a * b * q, -a], [-a * b * q, 1 - b * b * q, -b], [a, b, c]])
    T = np.array([0.0, 0.0, d])
    Rinv = R.T.copy()
    Tinv = Rinv @ (-T)
    return R, T, Rinv, Tinv


def xyz_compressed_bytes(valid, p3d, plane):
    """PovMesh.cpp:377-460: the bytes of mesh_cam.xyzC."""
    R, T, Rinv, Tinv = rt_from_plane(*plane)
    P = p3d[valid]
    n = P.shape[0]
    Q = np.stack([R[i, 0] * P[:, 0] + R[i, 1] * P[:, 1] + R[i, 2] * P[:, 2] + T[i] for i in range(3)], axis=1)
    mn, mx = Q.min(axis=0), Q.max(axis=0)
    scale = 65535.0 / (mx - mn)
    q16 = ((Q - mn) * scale).astype(np.uint16)  # C cast: truncation
    hdr = struct.pack("<I", n) + struct.pack("<3d", *scale) + struct.pack("<3d", *mn)
    hdr += struct.pack("<9d", *Rinv.reshape(-1)) + struct.pack("<3d", *Tinv)
    return hdr + q16.astype("<u2").tobytes()


def xyz_compressed_decode(buf):
    """The reader of gridding/wassgridsurface/wass_utils.py:22-35 (second format pin: matlab/load_camera_mesh.m:15-28)."""
    n = struct.unpack_from("<I", buf, 0)[0]
    scale = np.array(struct.unpack_from("<3d", buf, 4))
    mn = np.array(struct.unpack_from("<3d", buf, 28))
    Rinv = np.array(struct.unpack_from("<9d", buf, 52)).reshape(3, 3)
    Tinv = np.array(struct.unpack_from("<3d", buf, 124))
    q = np.frombuffer(buf, "<u2", count=3 * n, offset=148).reshape(n, 3).astype(np.float64)
    p_plane = q / scale + mn
    return (Rinv @ p_plane.T).T + Tinv


def load_camera_mesh_bytes(buf):
    """gridding/wassgridsurface/wass_utils.py:22-35 (load_camera_mesh), on bytes instead of a file name, same arithmetic:
    u16 -> float32 -> (promoted to float64) / scale + min, then Rinv @ p + Tinv.  Returns 3xN."""
    n = struct.unpack_from("<I", buf, 0)[0]
    limits = np.array(struct.unpack_from("<6d", buf, 4))
    Rinv = np.array(struct.unpack_from("<9d", buf, 52)).reshape(3, 3)
    Tinv = np.array(struct.unpack_from("<3d", buf, 124)).reshape(3, 1)
    data = np.frombuffer(buf, "<u2", count=3 * n, offset=148).reshape((3, n), order="F")
    m = data.astype(np.float32)
    m = m / np.expand_dims(limits[0:3], axis=1) + np.expand_dims(limits[3:6], axis=1)
    return Rinv @ m + Tinv


def align_on_sea_plane(mesh, plane, baseline=1.0):
    """wass_utils.py:38-68 (compute_sea_plane_RT, align_on_sea_plane_RT) and the `* baseline` of
    wassgridsurface.py:87,318.  mesh: 3xN."""
    a, b, c, d = plane
    q = (1 - c) / (a * a + b * b)
    R = np.array([[1 - a * a * q, -a * b * q, -a], [-a * b * q, 1 - b * b * q, -b], [a, b, c]])
    T = np.expand_dims(np.array([0, 0, d]), axis=1)
    out = R @ mesh + T
    out[2, :] *= -1.0
    return out * baseline


def rectified_calibration_identity(K, T, W, H):
    """Calibration of the synthetic rig of SURVEY.md §8d (already rectified: R=I, K0=K1, T along +x):
    what cv::stereoRectify(alpha=1) returns there -- R1=R2=I, P1=[K|0], P2=[K|(f*Tx,0,0)], ROIs (0,0,W-1,H-1)
    (pinned against cv2 in tests/test_oracle_pipeline.py)."""
    P1 = np.hstack([K, np.zeros((3, 1))])
    P2 = np.hstack([K, np.array([[K[0, 0] * T[0]], [0.0], [0.0]])])
    return dict(K0=K.copy(), K1=K.copy(), R=np.eye(3), T=np.asarray(T, np.float64), R1=np.eye(3), R2=np.eye(3),
                P1=P1, P2=P2, roi_left=(0, 0, W - 1, H - 1), roi_right=(0, 0, W - 1, H - 1))
