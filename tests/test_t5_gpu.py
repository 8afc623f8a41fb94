"""GPU parity of the UMT5 encoder schedule (realtime_video_b200/t5.py through the C ABI, bf16 like the server runs it)
against the reference T5Encoder goldens.  Tolerance: the reference's own bf16-vs-fp32 gap on this model is
rel-L2 7e-3 (tests/golden/make_t5_goldens.py); we require <= 2e-2 against both goldens.
"""
import pytest
import torch

from tests.golden_io import load_npz, rel_l2, weights

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag", ["a", "b"])
def test_t5_encoder_vs_reference_golden(tag):
    from realtime_video_b200.t5 import T5Encoder
    g = load_npz("t5_small.npz")
    m = T5Encoder(vocab=1000, dim=256, dim_attn=256, dim_ffn=512, num_heads=4, num_layers=2, num_buckets=32)
    m.load_state_dict(weights(g, torch.bfloat16), strict=True)
    m = m.to(device="cuda", dtype=torch.bfloat16).eval()
    y = m(g[f"{tag}/ids"].cuda(), g[f"{tag}/mask"].cuda())
    assert y.dtype == torch.bfloat16 and tuple(y.shape) == tuple(g[f"{tag}/fp32"].shape)
    for ref in ("fp32", "bf16"):
        r = rel_l2(y.float().cpu(), g[f"{tag}/{ref}"])
        assert r < 2e-2, f"{tag} vs {ref}: rel_l2={r:.3e}"
