"""Pin oracle/vae_oracle.py against outputs of the reference VAEDecoderWrapper
(tests/golden/make_vae_goldens.py).  CPU, fp32."""
import pytest
import torch

from oracle.vae_oracle import VAEDecoderOracle, synthetic_vae_params
from tests.golden_io import load_npz, rel_l2


@pytest.mark.parametrize("tag,sub", [("s8x12", 1), ("s16x24", 2)])
def test_decoder_oracle_matches_reference(tag, sub):
    g = load_npz("vae_small.npz")
    orc = VAEDecoderOracle(synthetic_vae_params(seed=0))
    cache = {}
    expect_frames = [9, 12, 4]
    with torch.no_grad():
        for call in range(3):
            px, cache = orc.forward(g[f"{tag}/z{call}"], cache)
            assert px.shape[1] == expect_frames[call]
            r = rel_l2(px[..., ::sub, ::sub], g[f"{tag}/px{call}"])
            assert r < 2e-5, f"{tag} call {call}: rel_l2={r:.3e}"


@pytest.mark.parametrize("tag", ["enc64x96", "enc128x192"])
def test_encoder_oracle_matches_reference(tag):
    from oracle.vae_oracle import VAEEncoderOracle
    g = load_npz("vae_small.npz")
    orc = VAEEncoderOracle(synthetic_vae_params(seed=0, encoder=True))
    with torch.no_grad():
        mu = orc.encode_first(g[f"{tag}/x"])
    assert mu.shape == g[f"{tag}/mu"].shape
    assert rel_l2(mu, g[f"{tag}/mu"]) < 2e-5
