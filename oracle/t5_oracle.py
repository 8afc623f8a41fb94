"""TEST INFRASTRUCTURE — CPU oracle for the UMT5 text encoder (not shipped, not measured; only tests/ import it).

Plain-torch restatement of the reference's ``T5Encoder.forward`` (wan/modules/t5.py:303-312) with
``shared_pos=False`` (UMT5): token embedding -> N x [ x += Attn(T5LayerNorm(x)) ; x += FFN(T5LayerNorm(x)) ] ->
T5LayerNorm.  T5LayerNorm = x * rsqrt(mean(x^2) + eps) * weight, no mean subtraction (:53-66); attention without
1/sqrt(d) scaling, additive bias = per-layer relative-position embedding (bidirectional buckets, :221-264) plus
the key padding mask filled with finfo.min, softmax in fp32 (:86-120); FFN = fc2(fc1(x) * gelu_tanh(gate(x)))
(:123-141); no biases anywhere.  Parameters: dict keyed like the reference state dict.
Pinned by tests/golden/make_t5_goldens.py (reference run on CPU) via tests/test_oracle_t5.py."""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def relative_position_bucket(lq: int, lk: int, num_buckets: int = 32, max_dist: int = 128) -> torch.Tensor:
    """t5.py:233-264, bidirectional: [lq, lk] bucket indices of rel = k - q."""
    rel = torch.arange(lk)[None, :] - torch.arange(lq)[:, None]
    nb = num_buckets // 2
    buckets = (rel > 0).long() * nb
    rel = rel.abs()
    max_exact = nb // 2
    large = max_exact + (torch.log(rel.float() / max_exact) / math.log(max_dist / max_exact) * (nb - max_exact)).long()
    large = torch.min(large, torch.full_like(large, nb - 1))
    return buckets + torch.where(rel < max_exact, rel, large)


def t5_layer_norm(x: torch.Tensor, weight: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    x = x * torch.rsqrt(x.float().pow(2).mean(dim=-1, keepdim=True) + eps)
    if weight.dtype in (torch.float16, torch.bfloat16):
        x = x.type_as(weight)
    return weight * x


def gelu_tanh(x: torch.Tensor) -> torch.Tensor:
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * torch.pow(x, 3.0))))


class T5EncoderOracle:
    def __init__(self, params, num_heads: int, num_buckets: int = 32):
        self.p, self.n, self.nb = params, num_heads, num_buckets
        self.layers = 1 + max(int(k.split(".")[1]) for k in params if k.startswith("blocks."))

    def forward(self, ids: torch.Tensor, mask: torch.Tensor | None = None) -> torch.Tensor:
        """ids [B, L] int64, mask [B, L] (1 = token) -> [B, L, dim]."""
        p, n = self.p, self.n
        x = F.embedding(ids, p["token_embedding.weight"])
        B, L, _ = x.shape
        buckets = relative_position_bucket(L, L, self.nb)
        for i in range(self.layers):
            pre = f"blocks.{i}."
            h = t5_layer_norm(x, p[pre + "norm1.weight"])
            q = (h @ p[pre + "attn.q.weight"].t()).view(B, L, n, -1)
            k = (h @ p[pre + "attn.k.weight"].t()).view(B, L, n, -1)
            v = (h @ p[pre + "attn.v.weight"].t()).view(B, L, n, -1)
            bias = x.new_zeros(B, n, L, L) + p[pre + "pos_embedding.embedding.weight"][buckets].permute(2, 0, 1)[None]
            if mask is not None:
                bias = bias.masked_fill(mask.view(B, 1, 1, L) == 0, torch.finfo(x.dtype).min)
            a = torch.einsum("binc,bjnc->bnij", q, k) + bias
            a = F.softmax(a.float(), dim=-1).type_as(a)
            o = torch.einsum("bnij,bjnc->binc", a, v).reshape(B, L, -1)
            x = x + o @ p[pre + "attn.o.weight"].t()
            h = t5_layer_norm(x, p[pre + "norm2.weight"])
            f = (h @ p[pre + "ffn.fc1.weight"].t()) * gelu_tanh(h @ p[pre + "ffn.gate.0.weight"].t())
            x = x + f @ p[pre + "ffn.fc2.weight"].t()
        return t5_layer_norm(x, p["norm.weight"])
