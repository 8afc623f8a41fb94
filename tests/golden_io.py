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


class ReplayRandn:
    """Feeds ``torch.randn(..., generator=...)`` calls from a recorded list (a golden's draws, in call order);
    calls without a generator pass through."""

    def __init__(self, draws):
        self.it, self.real = iter(draws), torch.randn

    def __enter__(self):
        def randn(*size, generator=None, **kw):
            if generator is None:
                return self.real(*size, **kw)
            shape = size[0] if len(size) == 1 and not isinstance(size[0], int) else size
            return next(self.it).to(device=kw.get("device", "cpu"), dtype=kw.get("dtype", torch.float32)).reshape(*shape)
        torch.randn = randn
        return self

    def __exit__(self, *exc):
        torch.randn = self.real
        return False
