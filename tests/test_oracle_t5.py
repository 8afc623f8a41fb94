"""Pin oracle/t5_oracle.py against the UNMODIFIED reference T5Encoder (tests/golden/make_t5_goldens.py), CPU."""
import pytest
import torch

from oracle.t5_oracle import T5EncoderOracle
from tests.golden_io import load_npz, rel_l2, weights


@pytest.mark.parametrize("tag", ["a", "b"])
def test_t5_oracle_matches_reference(tag):
    g = load_npz("t5_small.npz")
    o = T5EncoderOracle(weights(g, torch.float32), num_heads=4)
    y = o.forward(g[f"{tag}/ids"], g[f"{tag}/mask"])
    assert y.shape == g[f"{tag}/fp32"].shape
    assert rel_l2(y, g[f"{tag}/fp32"]) < 2e-5
    # the same restatement run in bf16 stays within the reference's own bf16-vs-fp32 gap (7e-3)
    ob = T5EncoderOracle(weights(g, torch.bfloat16), num_heads=4)
    yb = ob.forward(g[f"{tag}/ids"], g[f"{tag}/mask"]).float()
    assert rel_l2(yb, g[f"{tag}/bf16"]) < 5e-3
