"""Golden for the classic Self-Forcing loop: the UNMODIFIED reference
``CausalInferencePipeline.inference`` (pipeline/causal_inference.py:48-277) + ``WanVAEWrapper.
decode_to_pixel`` (utils/wan_wrapper.py:95-118) on CPU, fp32, small synthetic models.

    python tests/golden/make_pipeline_golden.py      # writes tests/golden/pipeline_small.npz

* DiT: the 2-layer fixture model of dit_small.npz (dim 256); latent 16x24 -> 96 tokens/frame, so
  the reference runs with its literal 1560 replaced by 96 ("patched-constant", ref_shim).
* VAE: reference WanVAE_ (dim 96) with ``synthetic_vae_params`` weights.
* The loop draws re-noising tensors with ``torch.randn_like`` (global RNG); they are recorded here
  in call order so the test can replay the identical values on the GPU.
* 2 blocks x 3 latent frames, warped 4-step schedule, context_noise 0 -> video [1, 21, 3, 128, 192].
"""
import sys
import types
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(HERE.parent.parent))
import ref_shim  # noqa: E402
from make_dit_goldens import build_model  # noqa: E402
from oracle.vae_oracle import synthetic_vae_params  # noqa: E402

if __name__ == "__main__":
    torch.set_num_threads(8)
    FS = 96
    ns = ref_shim.install(frame_seqlen=FS)
    ref_shim.patch_fp32_sdpa(ns)
    model = build_model(ns, torch.float32)
    W = ns.wan_wrapper.WanDiffusionWrapper
    gen = W.__new__(W)
    torch.nn.Module.__init__(gen)
    gen.model = model
    gen.uniform_timestep = False
    gen.scheduler = ns.scheduler.FlowMatchScheduler(shift=5.0, sigma_min=0.0, extra_one_step=True)
    gen.scheduler.set_timesteps(1000, training=True)
    gen.seq_len = 32760
    gen.post_init()
    V = ns.wan_wrapper.WanVAEWrapper
    vae = V.__new__(V)
    torch.nn.Module.__init__(vae)
    vae.mean = torch.tensor(ns.vae_block3.VAEDecoderWrapper().mean)
    vae.std = torch.tensor(ns.vae_block3.VAEDecoderWrapper().std)
    vae.model = ns.vae.WanVAE_(dim=96, z_dim=16, dim_mult=[1, 2, 4, 4], num_res_blocks=2, attn_scales=[],
                               temperal_downsample=[False, True, True], dropout=0.0)
    sd = synthetic_vae_params(seed=0)
    missing = vae.model.load_state_dict(sd, strict=False)
    assert not missing.unexpected_keys, missing.unexpected_keys
    vae.model = vae.model.float().eval()

    g = torch.Generator().manual_seed(21)
    ctx = torch.randn(1, 512, 128, generator=g)
    ctx[:, 20:] = 0
    text_encoder = lambda text_prompts: {"prompt_embeds": ctx}   # noqa: E731
    args = types.SimpleNamespace(denoising_step_list=[1000, 750, 500, 250], warp_denoising_step=True,
                                 num_frame_per_block=3, independent_first_frame=False, context_noise=0,
                                 model_kwargs={})
    pipe = ns.causal_inference.CausalInferencePipeline(args, "cpu", generator=gen, text_encoder=text_encoder, vae=vae)
    noise = torch.randn(1, 6, 16, 16, 24, generator=g)
    draws = []
    real = torch.randn_like

    def recording_randn_like(t, **kw):
        r = torch.randn(t.shape, generator=g, dtype=torch.float32).to(t.dtype)
        draws.append(r.clone())
        return r

    torch.randn_like = recording_randn_like
    try:
        with torch.no_grad():
            video, latents = pipe.inference(noise, ["x"], return_latents=True)
    finally:
        torch.randn_like = real
    out = {"ctx": ctx.numpy(), "noise": noise.numpy(), "latents": latents.numpy(),
           "video_sub": video[..., ::2, ::2].contiguous().numpy(),
           "steps": pipe.denoising_step_list.numpy()}
    for i, d in enumerate(draws):
        out[f"draw{i}"] = d.numpy()
    print("video", tuple(video.shape), "draws", len(draws), "steps", pipe.denoising_step_list.tolist())
    np.savez_compressed(HERE / "pipeline_small.npz", **out)
    print("pipeline_small.npz", sum(v.nbytes for v in out.values()) / 1e6, "MB raw")
