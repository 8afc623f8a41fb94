"""Executed in a SUBPROCESS by tests/test_reference_callers_cpu.py (the reference's release_server.py changes
process-wide torch state at import: grad mode, logging, dynamo limits; tests/ref_env.py patches torch.cuda).

    python tests/run_reference_callers.py server  [keep|reencode]
    python tests/run_reference_callers.py classic

``server``  : the UNMODIFIED ``release_server.GenerationSession`` (``__init__`` -> ``init_models`` ->
              ``generate_block_internal`` x 4: recompute_kv_cache, get_clean_context_frames, denoise loop, VAE decode,
              sliding window) driving the drop-in classes + the reference's own ``CausalInferencePipeline``; kernels
              replaced by the fp32 stand-ins of tests/cpu_ops_emulation.py (no GPU here).  Checked block by block
              (a) bit-equal against harness/server_loop.py — the stand-in bench/smoke/GPU tests use — and
              (b) against tests/golden/server_loop_small.npz, the same loop executed by the reference modules alone.
``classic`` : the UNMODIFIED ``CausalInferencePipeline.inference`` on the drop-in wrappers, bit-equal against
              harness/classic_loop.py and within tolerance of tests/golden/pipeline_small.npz.
Prints one JSON line; exit code 0 = all assertions held.
"""
import json
import sys
import types
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from tests import cpu_ops_emulation as emu  # noqa: E402
from tests import ref_env  # noqa: E402
from tests.golden_io import ReplayRandn as Replay, load_npz, rel_l2, weights  # noqa: E402

DIMS = dict(dim=256, ffn_dim=512, num_heads=2, num_layers=2, text_dim=128)


def use_standin_kernels():
    import realtime_video_b200.dit as dit
    import realtime_video_b200.vae as vae
    import realtime_video_b200.wan_wrapper as ww
    for mod in (dit, ww, vae):
        mod.ops = emu


def build_dropin_models(rs, g, width, height, text_encoder):
    """What ``load_all`` (release_server.py:250-313) produces, from synthetic weights instead of checkpoints: every
    class is obtained through the reference's import names."""
    from demo_utils.vae_block3 import VAEDecoderWrapper, VAEEncoderWrapper
    from pipeline import CausalInferencePipeline          # the reference's own pipeline class
    from utils.wan_wrapper import WanDiffusionWrapper

    import realtime_video_b200.vae as vae
    import realtime_video_b200.wan_wrapper as ww
    from realtime_video_b200.factory import synthetic_vae_params
    assert WanDiffusionWrapper is ww.WanDiffusionWrapper and VAEDecoderWrapper is vae.VAEDecoderWrapper
    assert CausalInferencePipeline.__module__ == "pipeline.causal_inference" and \
        "reference" in sys.modules["pipeline.causal_inference"].__file__

    def fresh():
        tr = WanDiffusionWrapper(model_name="synthetic", timestep_shift=5.0, is_causal=True, model_config=dict(DIMS))
        tr.model.load_state_dict(weights(g, torch.bfloat16), strict=False)
        tr = tr.to(dtype=torch.bfloat16).eval().requires_grad_(False)
        for blk in tr.model.blocks:                        # release_server.py:176-177
            blk.self_attn.fuse_projections()
        dec = VAEDecoderWrapper()
        dec.load_state_dict(synthetic_vae_params(seed=0), strict=False)
        dec = dec.to(dtype=torch.float16).eval().requires_grad_(False)
        enc = VAEEncoderWrapper()
        enc.load_state_dict(synthetic_vae_params(seed=0, encoder=True), strict=False)
        enc = enc.to(dtype=torch.float16).eval().requires_grad_(False)
        return tr, dec, enc

    tr, dec, enc = fresh()
    args = types.SimpleNamespace(denoising_step_list=[1000, 750, 500, 250], warp_denoising_step=True,
                                 num_frame_per_block=3, independent_first_frame=False, context_noise=0, model_kwargs={})
    pipe = CausalInferencePipeline(args, device="cpu", generator=tr, text_encoder=text_encoder, vae=dec)
    pipe.frame_seq_length = (height // 16) * (width // 16)      # other resolutions than 832x480
    return rs.Models(text_encoder, tr, pipe, enc, dec), fresh


def run_server(mode: str) -> dict:
    keep = mode == "keep"
    W, H, NB = 96, 64, 4
    rs = ref_env.setup(dropin=True)
    use_standin_kernels()
    g = load_npz("dit_small.npz")
    gold = load_npz("server_loop_small.npz")
    tag = "keep" if keep else "reenc"
    draws = [gold[f"{tag}/draw{i}"] for i in range(int(gold[f"{tag}/ndraws"]))]
    pe = gold["prompt_embeds"]

    if not keep:
        # release_server.py:574 re-encodes at a hard-coded 480x832; redirect that literal resize to the test
        # geometry (at 832x480 the bicubic resize is the identity), the same "patched-constant" device as the 1560
        import torch.nn.functional as F
        real_interp = F.interpolate

        def redirected(x, *a, **k):
            if k.get("size") is not None and tuple(k["size"]) == (480, 832):
                k["size"] = (H, W)
            return real_interp(x, *a, **k)
        F.interpolate = redirected

    class StaticText(torch.nn.Module):
        def forward(self, text_prompts):
            return {"prompt_embeds": pe.clone()}
    models, fresh = build_dropin_models(rs, g, W, H, StaticText())

    frames_cb = []
    params = rs.GenerateParams(prompt="x", width=W, height=H, seed=11, kv_cache_num_frames=3, num_blocks=NB,
                               num_denoising_steps=4, keep_first_frame=keep)
    config = types.SimpleNamespace(use_taehv=False)
    with Replay(draws):
        sess = rs.GenerationSession(params, config, frame_callback=lambda px, ids, ev: frames_cb.append(px.shape[1]),
                                    models=models)
        assert type(sess).__module__ == "release_server"
        real_px = [sess.generate_block_internal(models).clone() for _ in range(NB)]
    assert sess.generate_block_internal(models) is None and sess.block_idx == NB
    real_lat = sess.all_latents.clone()
    fs = (H // 16) * (W // 16)
    p = models.pipeline
    assert list(p.kv_cache1[0]["k"].shape) == [1, 6 * fs, 2, 128]
    assert [p.kv_cache1[0]["global_end_index"], p.kv_cache1[0]["local_end_index"]] == [6 * fs, 6 * fs]
    assert frames_cb == [6, 12, 12, 12]

    # (a) the harness stand-in, same seed / draws / weights: bit-equal
    import harness
    tr2, dec2, enc2 = fresh()
    hm = harness.build_models(tr2, vae_decoder=dec2, device="cpu", vae_encoder=enc2)
    hp = harness.GenerateParams(width=W, height=H, seed=11, kv_cache_num_frames=3, num_blocks=NB,
                                num_denoising_steps=4, keep_first_frame=keep)
    with Replay(draws):
        hs = harness.GenerationSession(hp, hm, prompt_embeds=pe, device="cpu")
        h_px = [hs.generate_block().clone() for _ in range(NB)]
    for b in range(NB):
        assert torch.equal(real_px[b], h_px[b]), f"harness differs from the real GenerationSession at block {b}"
    assert torch.equal(real_lat, hs.all_latents)

    # (b) the reference modules alone (golden): bf16 model / fp16 VAE on both sides, different arithmetic order
    out = {"mode": mode, "blocks": NB}
    r_lat = rel_l2(real_lat.float(), gold[f"{tag}/latents"].float())
    out["latents_rel_l2"] = r_lat
    assert r_lat < 1e-2, f"latents vs reference-executed golden: rel_l2={r_lat:.3e}"
    for b in range(NB):
        ref = gold[f"{tag}/px{b}_sub"].float()
        got = real_px[b][..., ::2, ::2]
        assert got.shape == ref.shape, (got.shape, ref.shape)
        mad = (got - ref).abs().mean().item()
        out[f"px{b}_mad"] = mad
        assert mad < 6e-3, f"block {b} pixels vs golden: mean|d|={mad:.3e}"
    return out


def run_classic() -> dict:
    rs = ref_env.setup(dropin=True)          # noqa: F841  (imports the reference tree with the finder installed)
    use_standin_kernels()
    from pipeline import CausalInferencePipeline
    from utils.wan_wrapper import WanDiffusionWrapper, WanVAEWrapper

    import harness
    from realtime_video_b200.factory import synthetic_vae_params
    g, gd = load_npz("pipeline_small.npz"), load_npz("dit_small.npz")
    ctx = g["ctx"].float()

    def build(cls):
        gen = WanDiffusionWrapper(model_name="synthetic", timestep_shift=5.0, is_causal=True, model_config=dict(DIMS))
        gen.model.load_state_dict(weights(gd, torch.float32), strict=False)
        gen = gen.float().eval()
        vae = WanVAEWrapper(load_pretrained=False)
        vae.model.load_state_dict(synthetic_vae_params(seed=0), strict=False)
        vae = vae.half().eval()
        args = types.SimpleNamespace(denoising_step_list=[1000, 750, 500, 250], warp_denoising_step=True,
                                     num_frame_per_block=3, independent_first_frame=False, context_noise=0,
                                     model_kwargs={})
        pipe = cls(args, "cpu", generator=gen, text_encoder=lambda text_prompts: {"prompt_embeds": ctx}, vae=vae)
        pipe.frame_seq_length = 96
        return pipe

    def run(pipe, twice=False):
        real = torch.randn_like
        outs = []
        for _ in range(2 if twice else 1):
            it = iter([g[f"draw{i}"] for i in range(6)])
            torch.randn_like = lambda t, **kw: next(it).to(device=t.device, dtype=t.dtype)
            try:
                with torch.no_grad():
                    outs.append(pipe.inference(g["noise"].float(), ["x"], return_latents=True))
            finally:
                torch.randn_like = real
        return outs

    ref_pipe = build(CausalInferencePipeline)
    assert "reference" in sys.modules[type(ref_pipe).__module__].__file__
    # two calls: the second goes through the reference's index-reset branch (:123-133, 1-element tensors)
    (v1, l1), (v2, l2) = run(ref_pipe, twice=True)
    assert torch.equal(l1, l2) and torch.equal(v1, v2)
    (hv, hl), (hv2, hl2) = run(build(harness.PipelineState), twice=True)
    assert torch.equal(l1, hl) and torch.equal(v1, hv) and torch.equal(hl, hl2)
    r = rel_l2(l1, g["latents"])
    mad = (v1[..., ::2, ::2] - g["video_sub"]).abs().mean().item()
    assert r < 1e-3 and mad < 5e-3, (r, mad)
    return {"mode": "classic", "latents_rel_l2": r, "video_mad": mad}


if __name__ == "__main__":
    torch.set_num_threads(8)
    what = sys.argv[1]
    res = run_server(sys.argv[2]) if what == "server" else run_classic()
    print(json.dumps(res))
