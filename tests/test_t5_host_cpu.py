"""Host schedule of realtime_video_b200/t5.py on the CPU (kernels replaced by the fp32 stand-ins) against the
goldens of the reference T5Encoder: per-head score / softmax / PV loop over column slices, V^T projection,
relative-position buckets and padding mask, residual epilogues, gated FFN.  fp32 on both sides: rel-L2 <= 1e-4."""
import pytest
import torch

from tests import cpu_ops_emulation as emu
from tests.golden_io import load_npz, rel_l2, weights


@pytest.fixture(autouse=True)
def cpu_ops(monkeypatch):
    import realtime_video_b200.t5 as t5
    monkeypatch.setattr(t5, "ops", emu)


def build(g, dtype=torch.float32):
    from realtime_video_b200.t5 import T5Encoder
    m = T5Encoder(vocab=1000, dim=256, dim_attn=256, dim_ffn=512, num_heads=4, num_layers=2, num_buckets=32)
    m.load_state_dict(weights(g, dtype), strict=True)                  # the reference's keys, nothing missing
    return m.to(dtype).eval()


@pytest.mark.parametrize("tag", ["a", "b"])
def test_t5_schedule_matches_reference(tag):
    g = load_npz("t5_small.npz")
    y = build(g)(g[f"{tag}/ids"], g[f"{tag}/mask"])
    assert y.shape == g[f"{tag}/fp32"].shape
    assert rel_l2(y, g[f"{tag}/fp32"]) < 1e-4


def test_no_mask_and_batch():
    g = load_npz("t5_small.npz")
    from oracle.t5_oracle import T5EncoderOracle
    m = build(g)
    ids = torch.cat([g["a/ids"], g["a/ids"].flip(1)])
    want = T5EncoderOracle(weights(g, torch.float32), num_heads=4).forward(ids, None)
    assert rel_l2(m(ids, None), want) < 1e-4


def test_umt5_xxl_dimensions_and_keys():
    from realtime_video_b200.t5 import UMT5_XXL, T5Encoder
    with torch.device("meta"):
        m = T5Encoder(**UMT5_XXL)
    sd = m.state_dict()
    assert sd["token_embedding.weight"].shape == (256384, 4096)
    assert sd["blocks.23.ffn.gate.0.weight"].shape == (10240, 4096)
    assert sd["blocks.0.pos_embedding.embedding.weight"].shape == (32, 64)
    assert sd["blocks.5.attn.o.weight"].shape == (4096, 4096) and "norm.weight" in sd
    assert len(sd) == 2 + 24 * 10


def test_drop_in_text_encoder_zeroes_padding_and_matches_oracle():
    """WanTextEncoder (utils/wan_wrapper.py:20-56) with an injected tokenizer: prompt_embeds [B, L, dim], rows past
    each prompt's length are zero, the valid rows equal the encoder output."""
    import realtime_video_b200.wan_wrapper as ww
    from oracle.t5_oracle import T5EncoderOracle
    g = load_npz("t5_small.npz")

    def tokenizer(prompts, return_mask=True, add_special_tokens=True):
        assert prompts == ["a prompt"] and return_mask and add_special_tokens
        return g["a/ids"], g["a/mask"]

    enc = ww.WanTextEncoder(model_config=dict(vocab=1000, dim=256, dim_attn=256, dim_ffn=512, num_heads=4,
                                              num_layers=2, num_buckets=32), tokenizer=tokenizer, device="cpu")
    enc.text_encoder.load_state_dict(weights(g, torch.float32), strict=True)
    out = enc(text_prompts=["a prompt"])["prompt_embeds"]
    assert out.shape == (1, 64, 256)
    valid = int(g["a/mask"].sum())
    assert float(out[0, valid:].abs().max()) == 0.0
    want = T5EncoderOracle(weights(g, torch.float32), num_heads=4).forward(g["a/ids"], g["a/mask"])
    assert rel_l2(out[0, :valid], want[0, :valid]) < 1e-4
    # without tokenizer files the failure is explicit
    enc.tokenizer = None
    with pytest.raises(FileNotFoundError, match="tokenizer"):
        enc(text_prompts=["x"])


def test_prompt_cleaning_matches_reference_rule():
    from realtime_video_b200.wan_wrapper import _PromptTokenizer
    assert _PromptTokenizer.clean("  a &amp;amp; b \n\t c  ") == "a & b c"
