"""Golden for the SERVER block loop: the UNMODIFIED ``release_server.GenerationSession``
(release_server.py:344-736) driving the UNMODIFIED reference modules — ``utils.wan_wrapper.WanDiffusionWrapper``
over ``CausalWanModel``, ``pipeline.causal_inference.CausalInferencePipeline``, ``demo_utils.vae_block3``
decoder / encoder wrappers — on the CPU in the build container.

    python tests/golden/make_server_loop_golden.py      # writes tests/golden/server_loop_small.npz

* 4 blocks, kv_cache_num_frames=3, 4 denoising steps, 96x64 px (latent 8x12, 24 tokens/frame): the context
  window starts sliding at block 2; stored for ``keep_first_frame=True`` ("keep") and for the reference's default
  first-frame re-encode ("reenc").
* DiT: the 2-layer fixture model of dit_small.npz in bf16 (the server's dtype; the KV cache and the latents are
  hard-coded bf16, release_server.py:397-404,549) with the reference's own bf16 SDPA fallbacks; VAE decoder /
  encoder fp16 like the server (:198,:208-217).
* CPU stand-ins, the same as for the other goldens (ref_shim): literal 1560 -> 24 ("patched-constant"), the flex
  branch evaluated as dense-mask SDPA with the reference's own ``get_sdpa_mask``; "reenc" additionally redirects the
  hard-coded 480x832 resize of release_server.py:574 to the test geometry (at 832x480 that resize is the identity).
* every ``torch.randn(..., generator=session.rnd)`` draw is recorded in call order so the tests replay the same values
  (a CUDA generator would produce different ones).
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
from realtime_video_b200.factory import synthetic_vae_params  # noqa: E402
from tests import ref_env  # noqa: E402

W, H, NB, FS = 96, 64, 4, 24
OUT = {}


def bf16_bits(t):
    return t.detach().contiguous().view(torch.int16).numpy()


@torch.no_grad()
def run(tag: str, keep: bool, pe: torch.Tensor):
    ns = ref_shim.install(frame_seqlen=FS)
    ref_shim.patch_flex_dense(ns)
    cm = ns.causal_model
    # recompute_kv_cache asks the model for a flex BlockMask (release_server.py:611-617); on the CPU the flex
    # call is the dense-mask SDPA stand-in, so hand it the reference's own dense mask of the same rule
    cm.CausalWanModel._prepare_blockwise_causal_attn_mask = staticmethod(
        lambda device, num_frames=21, frame_seqlen=FS, num_frame_per_block=1, local_attn_size=-1:
        cm.get_sdpa_mask("cpu", num_frames=num_frames, frame_seqlen=frame_seqlen,
                         num_frame_per_block=num_frame_per_block, local_attn_size=local_attn_size))
    rs = ref_env.setup(dropin=False)                       # unmodified release_server on the reference modules
    assert "reference" in sys.modules["utils.wan_wrapper"].__file__
    if not keep:
        import torch.nn.functional as F
        real_interp = F.interpolate

        def redirected(x, *a, **k):
            if k.get("size") is not None and tuple(k["size"]) == (480, 832):
                k["size"] = (H, W)
            return real_interp(x, *a, **k)
        F.interpolate = redirected

    model = build_model(ns, torch.bfloat16)
    Wr = ns.wan_wrapper.WanDiffusionWrapper
    tr = Wr.__new__(Wr)
    torch.nn.Module.__init__(tr)
    tr.model = model
    tr.uniform_timestep = False
    tr.scheduler = ns.scheduler.FlowMatchScheduler(shift=5.0, sigma_min=0.0, extra_one_step=True)
    tr.scheduler.set_timesteps(1000, training=True)
    tr.seq_len = 32760
    tr.post_init()
    for blk in tr.model.blocks:
        blk.self_attn.fuse_projections()

    dec = ns.vae_block3.VAEDecoderWrapper()
    dec.load_state_dict(synthetic_vae_params(seed=0), strict=False)
    dec = dec.to(dtype=torch.float16).eval().requires_grad_(False)
    vm = ns.vae.WanVAE_(dim=96, z_dim=16, dim_mult=[1, 2, 4, 4], num_res_blocks=2, attn_scales=[],
                        temperal_downsample=[False, True, True], dropout=0.0)
    vm.load_state_dict(synthetic_vae_params(seed=0, encoder=True), strict=False)
    enc = ns.vae_block3.VAEEncoderWrapper(types.SimpleNamespace(model=vm)).to(dtype=torch.float16).eval()

    args = types.SimpleNamespace(denoising_step_list=[1000, 750, 500, 250], warp_denoising_step=True,
                                 num_frame_per_block=3, independent_first_frame=False, context_noise=0, model_kwargs={})
    class StaticText(torch.nn.Module):
        def forward(self, text_prompts):
            return {"prompt_embeds": pe.clone()}
    text = StaticText()
    pipe = ns.causal_inference.CausalInferencePipeline(args, "cpu", generator=tr, text_encoder=text, vae=dec)
    assert pipe.frame_seq_length == FS

    models = rs.Models(text, tr, pipe, enc, dec)

    draws = []
    real_randn = torch.randn

    def recording_randn(*size, generator=None, **kw):
        r = real_randn(*size, generator=generator, **kw)
        if generator is not None:
            draws.append(r.clone())
        return r
    torch.randn = recording_randn
    try:
        params = rs.GenerateParams(prompt="x", width=W, height=H, seed=11, kv_cache_num_frames=3, num_blocks=NB,
                                   num_denoising_steps=4, keep_first_frame=keep)
        sess = rs.GenerationSession(params, types.SimpleNamespace(use_taehv=False),
                                    frame_callback=lambda *a: None, models=models)
        for b in range(NB):
            px = sess.generate_block_internal(models)
            OUT[f"{tag}/px{b}_sub"] = px[..., ::2, ::2].float().contiguous().numpy().astype(np.float16)
            print(tag, "block", b, tuple(px.shape), float(px.abs().mean()), flush=True)
    finally:
        torch.randn = real_randn
        if not keep:
            F.interpolate = real_interp
    OUT[f"{tag}/latents@bf16"] = bf16_bits(sess.all_latents)
    OUT[f"{tag}/ndraws"] = np.array(len(draws))
    for i, d in enumerate(draws):
        OUT[f"{tag}/draw{i}@bf16"] = bf16_bits(d)
    print(tag, "draws", len(draws), "steps", sess.denoising_step_list.tolist())
    ref_env.teardown()
    torch.set_grad_enabled(True)


if __name__ == "__main__":
    torch.set_num_threads(8)
    g = torch.Generator().manual_seed(5)
    pe = torch.randn(1, 512, 128, generator=g)
    pe[:, 20:] = 0
    pe = pe.to(torch.bfloat16)
    OUT["prompt_embeds@bf16"] = bf16_bits(pe)
    run("keep", True, pe)
    run("reenc", False, pe)
    np.savez_compressed(HERE / "server_loop_small.npz", **OUT)
    print("server_loop_small.npz", sum(v.nbytes for v in OUT.values()) / 1e6, "MB raw")
