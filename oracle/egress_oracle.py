"""CPU restatement of the reference's frame egress arithmetic — TEST INFRASTRUCTURE ONLY (imported by tests/;
never by the product path).

Reference: release_server.py:979-983 — after the device->host copy the fp32 frames [1, T, 3, H, W] get
``add_(1.0).mul_(0.5).clamp_(0.0, 1.0)``; release_server.py:973 hands each ``frames[0, idx]`` ([3, H, W] float)
to ``torchvision.transforms.functional.to_pil_image(..., "RGB")``, which for floating-point input does
``pic.mul(255).byte()`` and transposes CHW -> HWC (torchvision/transforms/functional.py, to_pil_image).
Everything is fp32 with one rounding per operation; ``.byte()`` truncates toward zero.

Pinned in tests/test_oracle_egress.py against those torch operations themselves (and against torchvision's
to_pil_image when torchvision + PIL are importable)."""
from __future__ import annotations

import numpy as np


def frames_to_rgb8(pixels: np.ndarray) -> np.ndarray:
    """pixels float32 [..., 3, H, W] -> uint8 [..., H, W, 3]."""
    x = np.asarray(pixels, dtype=np.float32)
    v = (x + np.float32(1.0)) * np.float32(0.5)          # two separately rounded fp32 operations
    v = np.clip(v, np.float32(0.0), np.float32(1.0))
    b = (v * np.float32(255.0)).astype(np.uint8)         # float -> uint8 truncation (values are in [0, 255])
    return np.moveaxis(b, -3, -1)
