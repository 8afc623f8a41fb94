"""``DO_COMPILE=true`` in the reference server wraps the transformer in ``torch.compile(module)`` and the VAE decoder in
``torch.compile(module, fullgraph=True)`` (release_server.py:753-755).  The drop-in modules must survive that: their
forwards are opaque frames (hand-scheduled C-ABI launches, nothing to trace), so the compiled modules run the same
code and return the same tensors.  Kernels replaced by the CPU stand-ins (host logic only)."""
import pytest
import torch

from tests import cpu_ops_emulation as emu
from tests.golden_io import load_npz, weights


@pytest.fixture(autouse=True)
def cpu_ops(monkeypatch):
    import realtime_video_b200.dit as dit
    import realtime_video_b200.vae as vae
    import realtime_video_b200.wan_wrapper as ww
    for mod in (dit, ww, vae):
        monkeypatch.setattr(mod, "ops", emu)


def test_compiled_transformer_wrapper_matches_eager():
    from realtime_video_b200.wan_wrapper import WanDiffusionWrapper
    g = load_npz("dit_small.npz")
    w = WanDiffusionWrapper(model_name="synthetic", timestep_shift=5.0, is_causal=True,
                            model_config=dict(dim=256, ffn_dim=512, num_heads=2, num_layers=2, text_dim=128))
    w.model.load_state_dict(weights(g, torch.float32), strict=False)
    w = w.float().eval()

    def caches():
        kv = [dict(k=torch.zeros(1, 6 * 96, 2, 128), v=torch.zeros(1, 6 * 96, 2, 128), global_end_index=0,
                   local_end_index=0) for _ in range(2)]
        ca = [dict(k=torch.zeros(1, 512, 2, 128), v=torch.zeros(1, 512, 2, 128), is_init=False) for _ in range(2)]
        return kv, ca

    lat = g["in/x0"].float().permute(1, 0, 2, 3)[None].contiguous()
    kw = dict(noisy_image_or_video=lat, conditional_dict={"prompt_embeds": g["in/ctx"].float()[None]},
              timestep=torch.ones(1, 3, dtype=torch.int64) * 750, current_start=0)
    kv, ca = caches()
    with torch.no_grad():
        f0, x0 = w(kv_cache=kv, crossattn_cache=ca, **kw)
        compiled = torch.compile(w)                                   # release_server.py:755
        kv2, ca2 = caches()
        f1, x1 = compiled(kv_cache=kv2, crossattn_cache=ca2, **kw)
    assert torch.equal(f0, f1) and torch.equal(x0, x1)
    assert kv2[0]["local_end_index"] == 3 * 96 and ca2[0]["is_init"]    # the cache dicts were mutated in place


def test_compiled_fullgraph_vae_decoder_matches_eager():
    from realtime_video_b200 import factory
    g = load_npz("vae_small.npz")
    z = g["s8x12/z0"].half()
    dec = factory.synthetic_vae_decoder(device="cpu")
    with torch.no_grad():
        px0, _ = dec(z, *([None] * 55))
        dec2 = factory.synthetic_vae_decoder(device="cpu")
        compiled = torch.compile(dec2, fullgraph=True)                # release_server.py:754
        px1, cache = compiled(z, *([None] * 55))
        assert torch.equal(px0, px1) and len(cache) == 55
        # second block of the stream: the cache list goes back in, like release_server.py:715
        z1 = g["s8x12/z1"].half()
        px0b, _ = dec(z1, *dec.engine.export_cache())
        px1b, cache = compiled(z1, *cache)
    assert px1b.shape == (1, 12, 3, 64, 96) and torch.equal(px0b, px1b)
