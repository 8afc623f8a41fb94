"""Launch the hot kernels at the BASELINE configs[1] shapes (for ncu captures)."""
import sys

import torch

sys.path.insert(0, ".")
from realtime_video_b200 import ops  # noqa: E402

which = sys.argv[1]
dev = "cuda"
torch.manual_seed(0)
L, D, FF, H = 4680, 5120, 13824, 40
if which == "gemm":
    a = torch.randn(L, D, device=dev, dtype=torch.bfloat16)
    w = torch.randn(FF, D, device=dev, dtype=torch.bfloat16) * 0.02
    b = torch.zeros(FF, device=dev, dtype=torch.bfloat16)
    o = torch.empty(L, FF, device=dev, dtype=torch.bfloat16)
    for _ in range(4):
        ops.gemm(a, w, b, epilogue=ops.EPI_BIAS_GELU, out=o)
elif which == "gemm1":   # N = 5120 (o / cross-attn q,o / ffn.2 shapes): the 1-CTA kernel
    a = torch.randn(L, FF, device=dev, dtype=torch.bfloat16)
    w = torch.randn(D, FF, device=dev, dtype=torch.bfloat16) * 0.02
    b = torch.zeros(D, device=dev, dtype=torch.bfloat16)
    o = torch.empty(L, D, device=dev, dtype=torch.bfloat16)
    for _ in range(4):
        ops.gemm(a, w, b, out=o)
elif which == "attn":
    q = torch.randn(L, D, device=dev, dtype=torch.bfloat16)
    k = torch.randn(2 * L, D, device=dev, dtype=torch.bfloat16)
    v = torch.randn(2 * L, D, device=dev, dtype=torch.bfloat16)
    o = torch.empty(L, D, device=dev, dtype=torch.bfloat16)
    for _ in range(4):
        ops.attention(q, k, v, heads=H, out=o)
elif which == "conv":
    from realtime_video_b200.vae import _prep_conv, _tile_for
    for (C, Hh, Ww, T) in [(96, 480, 832, 12), (192, 240, 416, 12)]:
        w = (torch.randn(C, C, 3, 3, 3) / (27 * C) ** 0.5).half()
        c = _prep_conv(w, torch.zeros(C).half(), torch.float16, dev)
        x = torch.randn(T + 2, Hh, Ww, C, device=dev, dtype=torch.float16)
        out = torch.empty(T, Hh, Ww, C, device=dev, dtype=torch.float16)
        nrm = torch.empty(T, Hh, Ww, C, device=dev, dtype=torch.float16)
        gamma = torch.ones(C, device=dev, dtype=torch.float16)
        for _ in range(3):
            ops.vae_conv(x, c.weight, c.bias, n=c.n, cout=c.cout, T=T, taps=(3, 3, 3), tile=_tile_for(Hh, Ww),
                         out_raw=out, out_norm=nrm, gamma=gamma, residual=out)
elif which == "flex":    # o-projection on a sequence-parallel shard (M = 585): wave-fitted runtime tile width
    a = torch.randn(585, D, device=dev, dtype=torch.bfloat16)
    w = torch.randn(D, D, device=dev, dtype=torch.bfloat16) * 0.02
    b = torch.zeros(D, device=dev, dtype=torch.bfloat16)
    o = torch.empty(585, D, device=dev, dtype=torch.bfloat16)
    for _ in range(4):
        ops.gemm(a, w, b, out=o)
elif which == "fp8":     # ffn.0 in e4m3
    from realtime_video_b200 import fp8
    a = torch.randn(L, D, device=dev, dtype=torch.bfloat16)
    w = torch.randn(FF, D, device=dev, dtype=torch.bfloat16) * 0.02
    wq, sw = fp8.quantize_weight(w)
    b = torch.zeros(FF, device=dev, dtype=torch.bfloat16)
    o = torch.empty(L, FF, device=dev, dtype=torch.bfloat16)
    for _ in range(4):
        ops.linear_fp8(a, wq, sw, b, epilogue=ops.EPI_BIAS_GELU, out=o)
elif which == "t5attn":
    q = torch.randn(512, 4096, device=dev, dtype=torch.bfloat16)
    k = torch.randn(512, 4096, device=dev, dtype=torch.bfloat16)
    v = torch.randn(512, 4096, device=dev, dtype=torch.bfloat16)
    bias = torch.randn(64, 1023, device=dev, dtype=torch.bfloat16)
    for _ in range(4):
        ops.t5_attention(q, k, v, bias, None, heads=64)
torch.cuda.synchronize()
print("done", which)
