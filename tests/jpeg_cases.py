"""Shared image generators / Pillow reference for the JPEG tests (oracle pinning, host emulation, GPU parity)."""
import io

import numpy as np


def pillow_jpeg(rgb: np.ndarray, quality: int = 90) -> bytes:
    """The reference's call (release_server.py:973) on an [H, W, 3] uint8 image."""
    from PIL import Image
    buf = io.BytesIO()
    Image.fromarray(np.ascontiguousarray(rgb), "RGB").save(buf, format="JPEG", quality=quality)
    return buf.getvalue()


def images(H: int, W: int, seed: int = 0):
    """name -> [H, W, 3] uint8: noise (long codes, many 0xFF bytes), smooth content (EOB-dominated), flat frames
    (DC difference 0, empty AC), extremes (largest coefficients, ZRL runs), low-amplitude noise."""
    rng = np.random.default_rng(seed + H * 7 + W)
    yy, xx = np.mgrid[0:H, 0:W]
    out = {
        "noise": (rng.random((H, W, 3)) * 256).astype(np.uint8),
        "gradient": np.stack([xx * 255 // W, yy * 255 // H, (xx + yy) * 255 // (H + W)], -1).astype(np.uint8),
        "black": np.zeros((H, W, 3), np.uint8),
        "white": np.full((H, W, 3), 255, np.uint8),
        "checker": np.stack([(((xx + yy) & 1) * 255).astype(np.uint8)] * 3, -1),
        "blocks": np.stack([(((xx // 8 + yy // 8) & 1) * 255).astype(np.uint8),
                            255 - (((xx // 8 + yy // 8) & 1) * 255).astype(np.uint8),
                            (((xx // 16) & 1) * 255).astype(np.uint8)], -1),
        "lownoise": (rng.random((H, W, 3)) * 40 + 100).astype(np.uint8),
        "smooth": np.clip(128 + 60 * np.sin(xx / 17.0)[..., None] + 40 * np.cos(yy / 11.0)[..., None]
                          + rng.normal(0, 3, (H, W, 3)), 0, 255).astype(np.uint8),
        "sparse": np.where(rng.random((H, W, 1)) < 0.02, 255, 16).astype(np.uint8).repeat(3, -1),
    }
    return {k: np.ascontiguousarray(v) for k, v in out.items()}


def frames_fp32(T: int, H: int, W: int, seed: int = 0) -> np.ndarray:
    """Decoder-like output float32 [T, 3, H, W]: smooth content + noise, some values outside [-1, 1]."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W]
    base = np.stack([np.sin(xx / 23.0 + t) * np.cos(yy / 19.0) for t in range(T)])[:, None]       # [T, 1, H, W]
    x = 0.8 * base + 0.15 * rng.normal(size=(T, 3, H, W)) + np.array([0.1, -0.2, 0.3])[None, :, None, None]
    x[:, :, : H // 4] *= 1.6                                                                      # clamped region
    return np.ascontiguousarray(x.astype(np.float32))
