"""The host executable's file codecs against cv2 on CPU: the PNG reader (`cv::imread(..., IMREAD_GRAYSCALE)` semantics for
every PNG flavour a camera pipeline produces: 1/2/4/8/16-bit grey, grey + alpha, RGB(A), palette) and the baseline JPEG
writer of the diagnostic images (decoded by cv2's libjpeg)."""
import os
import struct
import subprocess
import zlib
import numpy as np
import pytest
from helpers import ROOT

cv2 = pytest.importorskip("cv2")
HOST = os.path.join(ROOT, "wass_b200", "csrc", "host")


@pytest.fixture(scope="module")
def probe(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("io") / "io_probe")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I", HOST, os.path.join(ROOT, "tests", "io_probe.cpp"), os.path.join(HOST, "io.cpp"),
                    os.path.join(HOST, "jpeg.cpp"), "-lz", "-o", exe], check=True)
    return exe


def _chunk(t, d):
    return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xFFFFFFFF)


def _png(w, h, depth, ctype, rows, plte=None, filt=0, interlace=0):
    """rows: list of packed scanline bytes (already at the bit depth; for interlace=1 the scanlines of the seven Adam7
    passes in order); filt: PNG filter type byte, only 0 ("None") here -- real filters come from cv2.imwrite below."""
    raw = b"".join(bytes([filt]) + r for r in rows)
    out = b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, interlace))
    if plte is not None:
        out += _chunk(b"PLTE", bytes(plte))
    return out + _chunk(b"IDAT", zlib.compress(raw)) + _chunk(b"IEND", b"")


def _pack(vals, depth):
    """one scanline of sample values -> bytes, MSB first, padded to a byte"""
    bits = "".join(format(int(v), "0%db" % depth) for v in vals)
    bits += "0" * (-len(bits) % 8)
    return bytes(int(bits[i:i + 8], 2) for i in range(0, len(bits), 8))


def _read(probe, path, tmp_path):
    raw = str(tmp_path / "out.raw")
    r = subprocess.run([probe, "png", path, raw], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    rows, cols = (int(v) for v in r.stdout.split())
    return np.fromfile(raw, np.uint8).reshape(rows, cols)


@pytest.mark.parametrize("depth", [1, 2, 4])
@pytest.mark.parametrize("ctype", [0, 3])
def test_png_sub_byte_depths_match_cv2(probe, tmp_path, depth, ctype):
    rng = np.random.default_rng(depth * 10 + ctype)
    w, h = 37, 11                                     # a width that does not fill the last byte of a row
    vals = rng.integers(0, 1 << depth, (h, w))
    plte = rng.integers(0, 256, 3 << depth).tolist() if ctype == 3 else None
    p = str(tmp_path / "a.png")
    open(p, "wb").write(_png(w, h, depth, ctype, [_pack(r, depth) for r in vals], plte))
    ref = cv2.imread(p, cv2.IMREAD_GRAYSCALE)
    assert ref is not None and ref.shape == (h, w)
    assert np.array_equal(_read(probe, p, tmp_path), ref)


ADAM7 = [(0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2)]


@pytest.mark.parametrize("depth,ctype", [(8, 0), (1, 0), (4, 3), (8, 2), (16, 0)])
@pytest.mark.parametrize("shape", [(11, 37), (3, 2), (1, 1), (20, 9)])
def test_png_adam7_interlaced_matches_cv2(probe, tmp_path, depth, ctype, shape):
    rng = np.random.default_rng(depth * 100 + ctype * 10 + shape[0])
    h, w = shape
    ch = 3 if ctype == 2 else 1
    vals = rng.integers(0, 1 << min(depth, 8), (h, w, ch))
    plte = rng.integers(0, 256, 3 << depth).tolist() if ctype == 3 else None
    rows = []
    for x0, y0, dx, dy in ADAM7:
        sub = vals[y0::dy, x0::dx]
        if sub.shape[0] == 0 or sub.shape[1] == 0:
            continue
        for r in sub:
            if depth < 8:
                rows.append(_pack(r[:, 0], depth))
            elif depth == 8:
                rows.append(bytes(r.reshape(-1).tolist()))
            else:
                rows.append(b"".join(struct.pack(">H", int(v) * 257) for v in r[:, 0]))
    p = str(tmp_path / "i.png")
    open(p, "wb").write(_png(w, h, depth, ctype, rows, plte, interlace=1))
    ref = cv2.imread(p, cv2.IMREAD_GRAYSCALE)
    assert ref is not None and ref.shape == (h, w)
    assert np.array_equal(_read(probe, p, tmp_path), ref)


def test_png_flavours_written_by_cv2_match_cv2(probe, tmp_path):
    rng = np.random.default_rng(3)
    h, w = 53, 71
    smooth = (np.add.outer(np.arange(h), np.arange(w)) * 3 % 256).astype(np.uint8)      # makes the encoder pick real filters
    cases = {"g8": smooth, "g16": (smooth.astype(np.uint16) * 257 + 13), "rgb": np.dstack([smooth, smooth[::-1], rng.integers(0, 256, (h, w), dtype=np.uint8)]),
             "rgba": np.dstack([smooth, smooth[::-1], smooth.T[:h, :w] if smooth.T.shape == smooth.shape else smooth, rng.integers(0, 256, (h, w), dtype=np.uint8)]),
             "rgb16": np.dstack([smooth, smooth[::-1], smooth]).astype(np.uint16) * 200}
    for name, img in cases.items():
        p = str(tmp_path / (name + ".png"))
        assert cv2.imwrite(p, img)
        ref = cv2.imread(p, cv2.IMREAD_GRAYSCALE)
        got = _read(probe, p, tmp_path)
        assert got.shape == ref.shape, name
        assert np.array_equal(got, ref), (name, int(np.abs(got.astype(int) - ref).max()))


def test_png_palette_8bit_and_grey_alpha(probe, tmp_path):
    rng = np.random.default_rng(5)
    w, h = 19, 7
    idx = rng.integers(0, 256, (h, w))
    plte = rng.integers(0, 256, 768).tolist()
    p = str(tmp_path / "pal8.png")
    open(p, "wb").write(_png(w, h, 8, 3, [bytes(r.tolist()) for r in idx], plte))
    assert np.array_equal(_read(probe, p, tmp_path), cv2.imread(p, cv2.IMREAD_GRAYSCALE))
    ga = rng.integers(0, 256, (h, w, 2))
    ga[..., 1] = 255                                   # opaque: cv2 ignores alpha for IMREAD_GRAYSCALE
    p = str(tmp_path / "ga.png")
    open(p, "wb").write(_png(w, h, 8, 4, [bytes(r.reshape(-1).tolist()) for r in ga]))
    assert np.array_equal(_read(probe, p, tmp_path), cv2.imread(p, cv2.IMREAD_GRAYSCALE))


@pytest.mark.parametrize("shape", [(64, 96, 1), (61, 83, 1), (40, 56, 3), (37, 45, 3)])
def test_jpeg_writer_decodes_with_cv2(probe, tmp_path, shape):
    h, w, ch = shape
    y, x = np.mgrid[0:h, 0:w]
    base = (128 + 90 * np.sin(x / 9.0) * np.cos(y / 7.0)).astype(np.uint8)
    img = base[..., None] if ch == 1 else np.dstack([base, 255 - base, (x * 255 // w).astype(np.uint8)])
    raw, jpg = str(tmp_path / "in.raw"), str(tmp_path / "out.jpg")
    np.ascontiguousarray(img).tofile(raw)
    r = subprocess.run([probe, "jpeg", raw, str(h), str(w), str(ch), jpg], capture_output=True, text=True)
    assert r.returncode == 0
    dec = cv2.imread(jpg, cv2.IMREAD_UNCHANGED)
    assert dec is not None and dec.shape[:2] == (h, w)
    if ch == 1:
        assert dec.ndim == 2 and np.abs(dec.astype(int) - base).max() <= 3            # quality 95
    else:
        assert np.abs(dec[..., ::-1].astype(int) - img).max() <= 8                    # BGR -> RGB; colour conversion + quantiser
