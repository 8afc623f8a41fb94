"""GPU parity of the classic Self-Forcing loop (CausalInferencePipeline.inference +
WanVAEWrapper.decode_to_pixel) against a golden produced by the UNMODIFIED reference pipeline
(tests/golden/make_pipeline_golden.py).  The loop's torch.randn_like draws are replayed from the
fixture so both sides see identical re-noising.

Tolerance (stated): bf16 DiT through 2 blocks x (4 denoise + 1 context) passes with re-noising in
between, then a bf16 VAE, against the fp32 reference: latents rel-L2 <= 4e-2, video mean |d| <= 2e-2
(pixels in [0, 1])."""
import types

import pytest
import torch

from tests.golden_io import load_npz, rel_l2, weights

pytestmark = pytest.mark.gpu


def test_classic_inference_loop_vs_reference():
    from harness import PipelineState as CausalInferencePipeline
    from realtime_video_b200.wan_wrapper import WanDiffusionWrapper, WanVAEWrapper
    from realtime_video_b200.factory import synthetic_vae_params
    g = load_npz("pipeline_small.npz")
    gd = load_npz("dit_small.npz")
    gen = WanDiffusionWrapper(model_name="synthetic", timestep_shift=5.0, is_causal=True,
                              model_config=dict(dim=256, ffn_dim=512, num_heads=2, num_layers=2, text_dim=128))
    gen.model.load_state_dict(weights(gd, torch.bfloat16), strict=False)
    gen = gen.to(device="cuda", dtype=torch.bfloat16).eval()
    vae = WanVAEWrapper(load_pretrained=False)
    vae.model.load_state_dict(synthetic_vae_params(seed=0), strict=False)
    vae = vae.to(device="cuda", dtype=torch.bfloat16).eval()
    ctx = g["ctx"].cuda().to(torch.bfloat16)
    args = types.SimpleNamespace(denoising_step_list=[1000, 750, 500, 250], warp_denoising_step=True,
                                 num_frame_per_block=3, independent_first_frame=False, context_noise=0,
                                 model_kwargs={})
    pipe = CausalInferencePipeline(args, "cuda", generator=gen,
                                   text_encoder=lambda text_prompts: {"prompt_embeds": ctx}, vae=vae)
    pipe.frame_seq_length = 96          # (H/16)*(W/16); the reference's literal 1560 is 832x480
    assert torch.allclose(pipe.denoising_step_list.float(), g["steps"].float())
    draws = [g[f"draw{i}"] for i in range(6)]
    it = iter(draws)
    real = torch.randn_like
    torch.randn_like = lambda t, **kw: next(it).to(device=t.device, dtype=t.dtype)
    try:
        video, latents = pipe.inference(g["noise"].cuda().to(torch.bfloat16), ["x"], return_latents=True)
    finally:
        torch.randn_like = real
    assert video.shape == (1, 21, 3, 128, 192) and video.dtype == torch.float32
    assert float(video.min()) >= 0.0 and float(video.max()) <= 1.0
    r = rel_l2(latents, g["latents"])
    assert r < 4e-2, f"latents rel_l2={r:.3e}"
    mad = (video[..., ::2, ::2].cpu() - g["video_sub"]).abs().mean().item()
    assert mad < 2e-2, f"video mean|d|={mad:.3e}"
    assert pipe.kv_cache1[0]["k"].shape[1] == 32760     # the classic path's literal cache size (:289)
