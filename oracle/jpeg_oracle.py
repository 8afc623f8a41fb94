"""ctypes front end of ``oracle/jpeg_oracle.c`` — TEST INFRASTRUCTURE ONLY (imported by tests/, smoke() and
bench.py's CPU leg; never by the product path).

The C file restates libjpeg-turbo's baseline encoder for the parameters the reference passes
(``to_pil_image(frame, "RGB").save(io, format='JPEG', quality=90)``, release_server.py:973); it is pinned byte for
byte against Pillow in tests/test_oracle_jpeg.py.  Built on first use with gcc into ``oracle/_build/`` (git-ignored;
``__graft_entry__.build()`` builds it too)."""
from __future__ import annotations

import ctypes
import subprocess
from pathlib import Path

import numpy as np

_DIR = Path(__file__).resolve().parent
_SRC = _DIR / "jpeg_oracle.c"
_SO = _DIR / "_build" / "libjpeg_oracle.so"
_lib = None


def build(force: bool = False) -> Path:
    if force or not _SO.exists() or _SO.stat().st_mtime < _SRC.stat().st_mtime:
        _SO.parent.mkdir(exist_ok=True)
        tmp = _SO.with_suffix(".so.tmp")
        subprocess.run(["gcc", "-O2", "-std=c11", "-shared", "-fPIC", "-o", str(tmp), str(_SRC)], check=True)
        tmp.replace(_SO)
    return _SO


def _load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(str(build()))
        lib.jpeg_oracle_encode_rgb8.restype = ctypes.c_long
        lib.jpeg_oracle_encode_rgb8.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long]
        lib.jpeg_oracle_header.restype = ctypes.c_long
        lib.jpeg_oracle_header.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_long]
        _lib = lib
    return _lib


def encode_rgb8(rgb: np.ndarray, quality: int = 90, return_coefs: bool = False):
    """rgb uint8 [H, W, 3] (H, W multiples of 16) -> the bytes Pillow writes for
    ``Image.fromarray(rgb, 'RGB').save(io, 'JPEG', quality=quality)``; with ``return_coefs`` also the quantised
    coefficients int16 [blocks, 64] (zigzag order, scan order Y00 Y01 Y10 Y11 Cb Cr per 16x16 MCU)."""
    rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
    H, W, C = rgb.shape
    if C != 3 or H % 16 or W % 16:
        raise ValueError("jpeg oracle: need [H, W, 3] with H, W multiples of 16")
    scratch = np.empty(H * W * 3 // 2, np.uint8)
    coefs = np.zeros((H // 16 * (W // 16) * 6, 64), np.int16)
    cap = H * W * 3 + 4096
    out = np.empty(cap, np.uint8)
    n = _load().jpeg_oracle_encode_rgb8(rgb.ctypes.data, H, W, quality, scratch.ctypes.data, coefs.ctypes.data,
                                        out.ctypes.data, cap)
    if n <= 0:
        raise RuntimeError(f"jpeg oracle: output does not fit ({n})")
    data = out[:n].tobytes()
    return (data, coefs) if return_coefs else data


def header(H: int, W: int, quality: int = 90) -> bytes:
    """SOI .. SOS marker segments for an H x W image."""
    out = np.empty(1024, np.uint8)
    n = _load().jpeg_oracle_header(H, W, quality, out.ctypes.data, 1024)
    assert n > 0
    return out[:n].tobytes()


def frames_to_jpeg(pixels: np.ndarray, quality: int = 90) -> list:
    """Decoder output float32 [T, 3, H, W] in [-1, 1] -> T JPEG files: the reference's host-side
    normalisation + to_pil_image byte conversion (oracle/egress_oracle.py) followed by the encoder."""
    from oracle.egress_oracle import frames_to_rgb8
    rgb = frames_to_rgb8(pixels)
    return [encode_rgb8(f, quality) for f in rgb.reshape(-1, *rgb.shape[-3:])]
