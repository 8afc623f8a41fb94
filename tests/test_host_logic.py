"""Host-side logic that runs without a GPU: drop-in surface, state-dict contract, scheduler,
cache index algebra containers, denoising schedule, tile chooser."""
import types

import pytest
import torch

from tests.golden_io import load_npz, weights


def test_state_dict_contract_matches_reference_keys():
    """Keys of the fixture were written by the reference's own state_dict()."""
    from realtime_video_b200.dit import CausalWanModel
    g = load_npz("dit_small.npz")
    ref_keys = set(weights(g).keys())
    m = CausalWanModel(dim=256, ffn_dim=512, num_heads=2, num_layers=2, text_dim=128)
    assert set(m.state_dict().keys()) == ref_keys
    r = m.load_state_dict(weights(g), strict=True)
    assert not r.missing_keys and not r.unexpected_keys
    m.blocks[0].self_attn.fuse_projections()            # reference: extra to_qkv.* keys afterwards
    assert "blocks.0.self_attn.to_qkv.weight" in m.state_dict()
    assert m.blocks[0].self_attn.to_qkv.weight.shape == (768, 256)
    assert m.freqs.dtype == torch.complex128 and tuple(m.freqs.shape) == (1024, 64)


def test_vae_state_dict_contract():
    from realtime_video_b200.factory import synthetic_vae_params
    from realtime_video_b200.vae import VAEDecoderWrapper
    m = VAEDecoderWrapper()
    sd = synthetic_vae_params(0)
    mine = {k for k in m.state_dict().keys() if k not in ("mean", "std")}
    assert mine == set(sd.keys())
    n_conv3d = sum(1 for mod in m.decoder.modules() if mod.__class__.__name__ == "CausalConv3d")
    assert n_conv3d == 33                                # SURVEY.md §3.5 (count_conv3d)


def test_scheduler_matches_reference_tables():
    from realtime_video_b200.flow_match import FlowMatchSchedule as FlowMatchScheduler
    from harness import get_denoising_schedule
    g = load_npz("dit_small.npz")
    s = FlowMatchScheduler(shift=5.0, sigma_min=0.0, extra_one_step=True)
    s.set_timesteps(1000, training=True)
    assert torch.equal(s.sigmas, g["sched/sigmas"]) and torch.equal(s.timesteps, g["sched/timesteps"])
    clean, noise = g["in/x2"].permute(1, 0, 2, 3), g["in/x1"].permute(1, 0, 2, 3)
    out = s.add_noise(clean, noise, torch.full((3,), 750, dtype=torch.long))
    assert torch.equal(out, g["sched/add_noise_750"])
    zp = torch.cat((s.timesteps, torch.tensor([0], dtype=torch.float32)))
    for steps in (4, 5):
        assert torch.equal(get_denoising_schedule(zp, 1.0, steps), g[f"sched/steps{steps}"])


def test_pipeline_cache_allocation_and_reset():
    import harness
    from realtime_video_b200 import factory
    w = factory.synthetic_transformer("14B", device="cpu", dtype=torch.float32, num_layers=2, dim=256,
                                      ffn_dim=512, num_heads=2, text_dim=128)
    models = harness.build_models(w, device="cpu")
    p = models.pipeline
    assert p.frame_seq_length == 1560 and p.num_frame_per_block == 3
    p.local_attn_size = 6
    p._initialize_kv_cache(1, torch.bfloat16, "cpu")
    p._initialize_crossattn_cache(1, torch.bfloat16, "cpu")
    assert list(p.kv_cache1[0]["k"].shape) == [1, 6 * 1560, 2, 128]
    assert list(p.crossattn_cache[0]["k"].shape) == [1, 512, 2, 128]
    p.kv_cache1[0]["local_end_index"] = 99
    ptr = p.kv_cache1[0]["k"].data_ptr()
    p._initialize_kv_cache(1, torch.bfloat16, "cpu")     # same shape: re-initialise in place
    assert p.kv_cache1[0]["k"].data_ptr() == ptr and p.kv_cache1[0]["local_end_index"] == 0
    p.local_attn_size = -1
    p._initialize_kv_cache(1, torch.bfloat16, "cpu")
    assert p.kv_cache1[0]["k"].shape[1] == 32760          # reference default (causal_inference.py:289)
    # warped step list of the default config (release_server prints [1000, 937.5, 833.3, 625])
    assert torch.allclose(p.denoising_step_list, torch.tensor([1000.0, 937.5, 833.3333, 625.0]), atol=1e-3)


def test_block_mask_spec_and_tiles():
    from realtime_video_b200.dit import CausalWanModel
    from realtime_video_b200.vae import _tile_for
    spec = CausalWanModel._prepare_blockwise_causal_attn_mask("cuda", num_frames=5, frame_seqlen=1560,
                                                              num_frame_per_block=3, local_attn_size=-1)
    assert spec.block_len == 4680 and spec.window == 0
    for (h, w) in [(60, 104), (120, 208), (240, 416), (480, 832), (8, 12)]:
        tw, th = _tile_for(h, w)
        assert tw * th == 128
    assert _tile_for(120, 208) == (16, 8) and _tile_for(60, 104) == (8, 16)


def test_oracle_block_mask_equals_reference_rule():
    """block_causal_mask restates get_sdpa_mask / get_block_mask (causal_model.py:41-141)."""
    from oracle.dit_oracle import block_causal_mask
    L, bl = 20, 6
    m = block_causal_mask(L, L, bl)
    for q in range(L):
        end = (q // bl + 1) * bl
        for k in range(L):
            assert bool(m[q, k]) == (k < end or q == k)
