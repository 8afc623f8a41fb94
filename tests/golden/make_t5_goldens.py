"""Golden vectors for the UMT5 text encoder from the UNMODIFIED reference T5Encoder
(wan/modules/t5.py:267-313) on CPU.  Small configuration with the UMT5-XXL structure (head_dim 64, per-layer
relative position embeddings = shared_pos False, 32 buckets, gated tanh-GELU FFN, no biases):
vocab 1000, dim 256, dim_attn 256 (4 heads x 64), dim_ffn 512, 2 layers.  Weights: the reference's own
init_weights under torch.manual_seed(0) (norm weights re-drawn around 1 so they matter); they are stored in the
fixture.  Cases: 'a' 64 tokens of which 40 valid (padding mask), 'b' 512 tokens / 77 valid (the server's
text_len).  Outputs in fp32 and from the same model cast to bf16 (how the server runs it, release_server.py:141).
Run in the build container only:  python tests/golden/make_t5_goldens.py"""
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
sys.path.insert(0, str(HERE))
import ref_shim  # noqa: E402

CFG = dict(vocab=1000, dim=256, dim_attn=256, dim_ffn=512, num_heads=4, num_layers=2, num_buckets=32,
           shared_pos=False, dropout=0.1)


@torch.no_grad()
def main():
    import importlib
    ref_shim.install()
    t5 = importlib.import_module("wan.modules.t5")
    torch.manual_seed(0)
    m = t5.T5Encoder(**CFG).float().eval()
    g = torch.Generator().manual_seed(1)
    for n, p in m.named_parameters():
        if n.endswith("norm1.weight") or n.endswith("norm2.weight") or n == "norm.weight":
            p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
    out = {}
    for k, v in m.state_dict().items():
        out["w/" + k] = v.numpy()
    mb = t5.T5Encoder(**CFG).eval()
    mb.load_state_dict(m.state_dict())
    mb = mb.to(torch.bfloat16)
    for tag, L, valid in (("a", 64, 40), ("b", 512, 77)):
        ids = torch.randint(1, CFG["vocab"], (1, L), generator=g)
        mask = torch.zeros(1, L, dtype=torch.long)
        mask[:, :valid] = 1
        ids[:, valid:] = 0
        y = m(ids, mask)
        yb = mb(ids, mask)
        out[f"{tag}/ids"], out[f"{tag}/mask"] = ids.numpy(), mask.numpy()
        out[f"{tag}/fp32"] = y.numpy()
        out[f"{tag}/bf16"] = yb.float().numpy()
        print(tag, tuple(y.shape), float(y.abs().mean()), "bf16 gap", float((yb.float() - y).norm() / y.norm()))
    np.savez_compressed(HERE / "t5_small.npz", **out)
    print("t5_small.npz", sum(v.nbytes for v in out.values()) / 1e6, "MB raw")


if __name__ == "__main__":
    main()
