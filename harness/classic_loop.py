"""TEST / BENCH INFRASTRUCTURE — stand-in for the reference's ``CausalInferencePipeline``
(pipeline/causal_inference.py:9-339) on boxes where the reference checkout does not exist (the GPU box).

In a real deployment the reference's OWN, unmodified ``CausalInferencePipeline`` runs on top of the drop-in
wrappers (realtime_video_b200/dropin); ``tests/test_reference_callers_cpu.py`` executes exactly that and
asserts bit-equality with this stand-in, so results obtained with the stand-in on the GPU carry over.
Only what the hot path needs is here: the cache allocators the server calls, and the text-to-video branch
of the classic loop (no initial latent, no profiling, no low-memory offload).
"""
from __future__ import annotations

from typing import List

import torch


class PipelineState(torch.nn.Module):
    """Attribute-compatible with the reference class for the callers in harness/server_loop.py."""

    def __init__(self, args, device, generator, text_encoder, vae):
        super().__init__()
        self.generator, self.text_encoder, self.vae, self.args = generator, text_encoder, vae, args
        self.scheduler = generator.get_scheduler()
        steps = torch.tensor(args.denoising_step_list, dtype=torch.long)
        if args.warp_denoising_step:         # (:31-34) look the step list up in the shifted schedule
            table = torch.cat((self.scheduler.timesteps.cpu(), torch.tensor([0], dtype=torch.float32)))
            steps = table[1000 - steps]
        self.denoising_step_list = steps
        self.num_transformer_blocks = len(generator.model.blocks)
        self.frame_seq_length = 1560         # (:35) the reference's literal; callers at other sizes overwrite it
        self.kv_cache1 = None
        self.num_frame_per_block = getattr(args, "num_frame_per_block", 1)
        self.independent_first_frame = args.independent_first_frame
        self.local_attn_size = generator.model.local_attn_size
        if self.num_frame_per_block > 1:
            generator.model.num_frame_per_block = self.num_frame_per_block

    # -- (:279-339): allocate once, afterwards zero in place and reset the indices --------------------------
    def _cache_heads(self) -> int:
        # a sequence-parallel rank keeps only its own heads (realtime_video_b200/parallel.py)
        return getattr(self.generator.model, "kv_cache_heads", self.generator.model.config.num_heads)

    def _initialize_kv_cache(self, batch_size, dtype, device):
        cfg = self.generator.model.config
        rows = 32760 if self.local_attn_size == -1 else self.local_attn_size * self.frame_seq_length
        shape = [batch_size, rows, self._cache_heads(), cfg.dim // cfg.num_heads]
        if self.kv_cache1 and list(self.kv_cache1[0]["k"].shape) == shape:
            for c in self.kv_cache1:
                c["k"].zero_()
                c["v"].zero_()
                c["global_end_index"] = c["local_end_index"] = 0
            return
        sp = getattr(self.generator.model, "sp", None)
        if sp is not None and sp.p2p:
            # multi-GPU single-stream mode: the caches live in the rank's symmetric allocation so that the peers'
            # kernels can store K / V rows straight into them (realtime_video_b200/parallel.py)
            kv = sp.alloc_kv_cache(self.num_transformer_blocks, shape, dtype, device)
            if kv is not None:
                self.kv_cache1 = [dict(k=k, v=v, global_end_index=0, local_end_index=0) for k, v in kv]
                return
            # symmetric memory unavailable: sp has switched to the collective exchange on every rank
        self.kv_cache1 = [dict(k=torch.zeros(shape, dtype=dtype, device=device),
                               v=torch.zeros(shape, dtype=dtype, device=device),
                               global_end_index=0, local_end_index=0) for _ in range(self.num_transformer_blocks)]

    def _initialize_crossattn_cache(self, batch_size, dtype, device):
        cfg = self.generator.model.config
        shape = [batch_size, 512, cfg.num_heads, cfg.dim // cfg.num_heads]
        old = getattr(self, "crossattn_cache", None)
        if old and list(old[0]["k"].shape) == shape:
            for c in old:
                c["k"].zero_()
                c["v"].zero_()
                c["is_init"] = False
            return
        self.crossattn_cache = [dict(k=torch.zeros(shape, dtype=dtype, device=device),
                                     v=torch.zeros(shape, dtype=dtype, device=device), is_init=False)
                                for _ in range(self.num_transformer_blocks)]

    # -- (:48-277), text-to-video branch ----------------------------------------------------------------------
    @torch.no_grad()
    def inference(self, noise: torch.Tensor, text_prompts: List[str], return_latents: bool = False):
        B, F, C, H, W = noise.shape
        nf = self.num_frame_per_block
        assert F % nf == 0 and not self.independent_first_frame
        cond = self.text_encoder(text_prompts=text_prompts)
        latents = torch.zeros_like(noise)
        if self.kv_cache1 is None:
            self._initialize_kv_cache(B, noise.dtype, noise.device)
            self._initialize_crossattn_cache(B, noise.dtype, noise.device)
        else:                                 # (:123-133) index reset with 1-element tensors, as the reference
            for c in self.crossattn_cache:
                c["is_init"] = False
            for c in self.kv_cache1:
                c["global_end_index"] = torch.tensor([0], dtype=torch.long, device=noise.device)
                c["local_end_index"] = torch.tensor([0], dtype=torch.long, device=noise.device)
        caches = dict(conditional_dict=cond, kv_cache=self.kv_cache1, crossattn_cache=self.crossattn_cache)
        steps = self.denoising_step_list
        for f0 in range(0, F, nf):
            x, start = noise[:, f0:f0 + nf], f0 * self.frame_seq_length
            for i, t in enumerate(steps):
                ts = torch.ones([B, nf], device=noise.device, dtype=torch.int64) * t
                _, x0 = self.generator(noisy_image_or_video=x, timestep=ts, current_start=start, **caches)
                if i + 1 < len(steps):        # (:205-213) re-noise to the next step with the global RNG
                    flat = x0.flatten(0, 1)
                    nxt = steps[i + 1] * torch.ones([B * nf], device=noise.device, dtype=torch.long)
                    x = self.scheduler.add_noise(flat, torch.randn_like(flat), nxt).unflatten(0, x0.shape[:2])
            latents[:, f0:f0 + nf] = x0
            # (:227-236) one more pass at the context-noise timestep leaves the CLEAN K/V in the cache
            self.generator(noisy_image_or_video=x0, timestep=torch.ones_like(ts) * self.args.context_noise,
                           current_start=start, **caches)
        video = (self.vae.decode_to_pixel(latents, use_cache=False) * 0.5 + 0.5).clamp(0, 1)
        return (video, latents) if return_latents else video
