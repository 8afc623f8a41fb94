"""Load the committed golden fixtures (tests/golden/*.npz)."""
from pathlib import Path

import numpy as np
import torch

GOLDEN = Path(__file__).resolve().parent / "golden"


def load_npz(name: str) -> dict:
    out = {}
    with np.load(GOLDEN / name) as z:
        for k in z.files:
            a = z[k]
            if k.endswith("@bf16"):
                out[k[:-5]] = torch.from_numpy(a.copy()).view(torch.bfloat16)
            else:
                out[k] = torch.from_numpy(a.copy())
    return out


def weights(g: dict, dtype=torch.float32, prefix: str = "w/", device="cpu") -> dict:
    return {k[len(prefix):]: v.to(device=device, dtype=dtype) for k, v in g.items() if k.startswith(prefix)}


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()
