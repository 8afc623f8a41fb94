"""TEST INFRASTRUCTURE — CPU oracle for the FP8 linear (not shipped; only tests/ import it).  Parity PARTLY pinned:
the scaled matmul + bias + bf16 rounding is pinned to ``torch._scaled_mm`` on the CPU — the kernel torchao's Float8
linear dispatches to (tests/test_fp8_cpu.py::test_oracle_linear_matches_torch_scaled_mm); the per-tensor scale formula
is "parity unpinned": torchao is not in the image and the reference ships no FP8 vectors, so it restates the PUBLISHED
algorithm of
``Float8DynamicActivationFloat8WeightConfig(granularity=PerTensor())`` (torchao float8 inference: per-tensor
``scale = finfo(e4m3).max / clamp(amax, 1e-12)``; ``x_q = (x * scale).clamp(+-448).to(float8_e4m3fn)``;
``torch._scaled_mm(x_q, w_q.t(), scale_a = 1/scale_x, scale_b = 1/scale_w, bias, out_dtype = bf16)`` with fp32
accumulation) as it is applied at release_server.py:179-182.  torch's own float8_e4m3fn cast (round to nearest even,
saturating after the clamp) supplies the rounding."""
from __future__ import annotations

import torch

E4M3_MAX = 448.0


def quantize_per_tensor(x: torch.Tensor):
    """-> (e4m3 tensor, dequantisation scale = amax / 448 as fp32 scalar tensor)."""
    amax = x.abs().max().float().clamp(min=1e-12)
    q = (x.float() * (E4M3_MAX / amax)).clamp(-E4M3_MAX, E4M3_MAX).to(torch.float8_e4m3fn)
    return q, amax / E4M3_MAX


def linear_fp8(x: torch.Tensor, w: torch.Tensor, bias: torch.Tensor | None = None) -> torch.Tensor:
    """bf16(x_q @ w_q^T * s_x * s_w + bias) with x, w given in bf16 / fp32 (quantised here)."""
    xq, sx = quantize_per_tensor(x)
    wq, sw = quantize_per_tensor(w)
    y = (xq.float() @ wq.float().t()) * (sx * sw)
    if bias is not None:
        y = y + bias.float()
    return y.to(torch.bfloat16)
