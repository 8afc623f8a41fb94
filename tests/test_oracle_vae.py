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


def test_streaming_encoder_oracle_matches_reference():
    """VAEEncoderOracle.forward (chunks of 1, 4, 4 frames cold; 4, 4 / 4 frames with stream=True on the carried
    cache) against the UNMODIFIED reference VAEEncoderWrapper (tests/golden/make_vae_encoder_stream_goldens.py):
    fp32 on both sides, rel-L2 <= 2e-5."""
    from oracle.vae_oracle import VAEEncoderOracle
    g = load_npz("vae_encoder_stream.npz")
    o = VAEEncoderOracle(synthetic_vae_params(seed=0, encoder=True))
    cache = {}
    for tag, stream in (("cold9_64x96", False), ("stream8_64x96", True), ("stream4_64x96", True)):
        mu, cache = o.forward(g[f"{tag}/x"], cache, stream=stream)
        assert mu.shape == g[f"{tag}/mu"].shape
        assert rel_l2(mu, g[f"{tag}/mu"]) < 2e-5, tag
    assert len([c for c in cache.values() if c is not None]) == int(g["n_cache_slots"][0])
    mu, _ = o.forward(g["cold5_48x80/x"], {}, stream=False)
    assert rel_l2(mu, g["cold5_48x80/mu"]) < 2e-5
    # the first-chunk shortcut used by the server's first-frame re-encode is the same function
    one = g["cold9_64x96/x"][:, :, :1]
    assert torch.equal(o.encode_first(one), o.forward(one, {}, stream=False)[0])
