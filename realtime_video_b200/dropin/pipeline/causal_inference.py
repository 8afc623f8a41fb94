"""Drop-in for ``pipeline/causal_inference.py`` (reference pipeline/causal_inference.py:9-339).

State holder for the server loop (release_server.py:542-736 reads/writes ``kv_cache1``,
``crossattn_cache``, ``frame_seq_length``, ``local_attn_size``, ``scheduler`` ...) and the
classic Self-Forcing loop ``inference``.  Differences from the reference, both observable-
behaviour preserving:
  * ``frame_seq_length`` defaults to 1560 (832x480) but follows the latent size once a call
    shows another resolution (the reference hard-codes 1560);
  * re-initialising an existing KV cache resets the index fields only: rows at or beyond
    ``local_end_index`` are never read (causal_model.py:388), so the reference's 7.7 GB
    ``zero_()`` per block is skipped.  ``zero_kv_on_reset=True`` restores the memset.
"""
from __future__ import annotations

from typing import List, Optional

import torch

from realtime_video_b200.dropin.utils.wan_wrapper import (WanDiffusionWrapper, WanTextEncoder,
                                                          WanVAEWrapper)


class CausalInferencePipeline(torch.nn.Module):
    def __init__(self, args, device, generator=None, text_encoder=None, vae=None):
        super().__init__()
        self.generator = WanDiffusionWrapper(**getattr(args, "model_kwargs", {}), is_causal=True) \
            if generator is None else generator
        self.text_encoder = WanTextEncoder() if text_encoder is None else text_encoder
        self.vae = WanVAEWrapper() if vae is None else vae

        self.scheduler = self.generator.get_scheduler()
        self.denoising_step_list = torch.tensor(args.denoising_step_list, dtype=torch.long)
        if args.warp_denoising_step:
            timesteps = torch.cat((self.scheduler.timesteps.cpu(), torch.tensor([0], dtype=torch.float32)))
            self.denoising_step_list = timesteps[1000 - self.denoising_step_list]

        self.num_transformer_blocks = len(self.generator.model.blocks)
        self.frame_seq_length = 1560
        self.kv_cache1 = None
        self.args = args
        self.num_frame_per_block = getattr(args, "num_frame_per_block", 1)
        self.independent_first_frame = args.independent_first_frame
        self.local_attn_size = self.generator.model.local_attn_size
        self.zero_kv_on_reset = False
        if self.num_frame_per_block > 1:
            self.generator.model.num_frame_per_block = self.num_frame_per_block

    # -- pipeline/causal_inference.py:48-277 ---------------------------------------------------
    @torch.no_grad()
    def inference(self, noise: torch.Tensor, text_prompts: List[str],
                  initial_latent: Optional[torch.Tensor] = None, return_latents: bool = False,
                  profile: bool = False, low_memory: bool = False):
        batch_size, num_frames, num_channels, height, width = noise.shape
        self.frame_seq_length = (height // 2) * (width // 2)
        if not self.independent_first_frame or (self.independent_first_frame and initial_latent is not None):
            assert num_frames % self.num_frame_per_block == 0
            num_blocks = num_frames // self.num_frame_per_block
        else:
            assert (num_frames - 1) % self.num_frame_per_block == 0
            num_blocks = (num_frames - 1) // self.num_frame_per_block
        num_input_frames = initial_latent.shape[1] if initial_latent is not None else 0
        num_output_frames = num_frames + num_input_frames
        conditional_dict = self.text_encoder(text_prompts=text_prompts)
        output = torch.zeros([batch_size, num_output_frames, num_channels, height, width],
                             device=noise.device, dtype=noise.dtype)
        ev = {}
        if profile:
            for k in ("init_s", "init_e", "diff_s", "diff_e", "vae_s", "vae_e"):
                ev[k] = torch.cuda.Event(enable_timing=True)
            block_times, blk_s, blk_e = [], torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev["init_s"].record()

        # Step 1: caches (:112-133)
        if self.kv_cache1 is None:
            self._initialize_kv_cache(batch_size=batch_size, dtype=noise.dtype, device=noise.device)
            self._initialize_crossattn_cache(batch_size=batch_size, dtype=noise.dtype, device=noise.device)
        else:
            for i in range(self.num_transformer_blocks):
                self.crossattn_cache[i]["is_init"] = False
            for c in self.kv_cache1:
                c["global_end_index"] = 0
                c["local_end_index"] = 0

        # Step 2: context frames (:135-170)
        current_start_frame = 0
        gen = dict(conditional_dict=conditional_dict, kv_cache=self.kv_cache1,
                   crossattn_cache=self.crossattn_cache)
        if initial_latent is not None:
            timestep = torch.zeros([batch_size, 1], device=noise.device, dtype=torch.int64)
            if self.independent_first_frame:
                assert (num_input_frames - 1) % self.num_frame_per_block == 0
                num_input_blocks = (num_input_frames - 1) // self.num_frame_per_block
                output[:, :1] = initial_latent[:, :1]
                self.generator(noisy_image_or_video=initial_latent[:, :1], timestep=timestep * 0,
                               current_start=current_start_frame * self.frame_seq_length, **gen)
                current_start_frame += 1
            else:
                assert num_input_frames % self.num_frame_per_block == 0
                num_input_blocks = num_input_frames // self.num_frame_per_block
            for _ in range(num_input_blocks):
                ref = initial_latent[:, current_start_frame:current_start_frame + self.num_frame_per_block]
                output[:, current_start_frame:current_start_frame + self.num_frame_per_block] = ref
                ts = torch.zeros([batch_size, self.num_frame_per_block], device=noise.device, dtype=torch.int64)
                self.generator(noisy_image_or_video=ref, timestep=ts,
                               current_start=current_start_frame * self.frame_seq_length, **gen)
                current_start_frame += self.num_frame_per_block
        if profile:
            ev["init_e"].record()
            torch.cuda.synchronize()
            ev["diff_s"].record()

        # Step 3: temporal denoising loop (:177-245)
        all_num_frames = [self.num_frame_per_block] * num_blocks
        if self.independent_first_frame and initial_latent is None:
            all_num_frames = [1] + all_num_frames
        for current_num_frames in all_num_frames:
            if profile:
                blk_s.record()
            noisy_input = noise[:, current_start_frame - num_input_frames:
                                current_start_frame + current_num_frames - num_input_frames]
            for index, current_timestep in enumerate(self.denoising_step_list):
                timestep = torch.ones([batch_size, current_num_frames], device=noise.device,
                                      dtype=torch.int64) * current_timestep
                _, denoised_pred = self.generator(
                    noisy_image_or_video=noisy_input, timestep=timestep,
                    current_start=current_start_frame * self.frame_seq_length, **gen)
                if index < len(self.denoising_step_list) - 1:
                    next_timestep = self.denoising_step_list[index + 1]
                    noisy_input = self.scheduler.add_noise(
                        denoised_pred.flatten(0, 1), torch.randn_like(denoised_pred.flatten(0, 1)),
                        next_timestep * torch.ones([batch_size * current_num_frames], device=noise.device,
                                                   dtype=torch.long)).unflatten(0, denoised_pred.shape[:2])
            output[:, current_start_frame:current_start_frame + current_num_frames] = denoised_pred
            # rerun at the context-noise timestep to store clean K/V (:227-236)
            context_timestep = torch.ones_like(timestep) * self.args.context_noise
            self.generator(noisy_image_or_video=denoised_pred, timestep=context_timestep,
                           current_start=current_start_frame * self.frame_seq_length, **gen)
            if profile:
                blk_e.record()
                torch.cuda.synchronize()
                block_times.append(blk_s.elapsed_time(blk_e))
            current_start_frame += current_num_frames
        if profile:
            ev["diff_e"].record()
            torch.cuda.synchronize()
            ev["vae_s"].record()

        # Step 4: decode (:255-257)
        video = self.vae.decode_to_pixel(output, use_cache=False)
        video = (video * 0.5 + 0.5).clamp(0, 1)
        if profile:
            ev["vae_e"].record()
            torch.cuda.synchronize()
            init_t = ev["init_s"].elapsed_time(ev["init_e"])
            diff_t = ev["diff_s"].elapsed_time(ev["diff_e"])
            vae_t = ev["vae_s"].elapsed_time(ev["vae_e"])
            total = init_t + diff_t + vae_t
            print("Profiling results:")
            print(f"  - Initialization/caching time: {init_t:.2f} ms ({100 * init_t / total:.2f}%)")
            print(f"  - Diffusion generation time: {diff_t:.2f} ms ({100 * diff_t / total:.2f}%)")
            for i, bt in enumerate(block_times):
                print(f"    - Block {i} generation time: {bt:.2f} ms ({100 * bt / diff_t:.2f}% of diffusion)")
            print(f"  - VAE decoding time: {vae_t:.2f} ms ({100 * vae_t / total:.2f}%)")
            print(f"  - Total time: {total:.2f} ms")
        return (video, output) if return_latents else video

    # -- pipeline/causal_inference.py:279-339 --------------------------------------------------
    def _initialize_kv_cache(self, batch_size, dtype, device):
        if self.local_attn_size != -1:
            kv_cache_size = self.local_attn_size * self.frame_seq_length
        else:
            kv_cache_size = 21 * self.frame_seq_length      # 32760 at 1560 tokens/frame (:289)
        head_dim = self.generator.model.config.dim // self.generator.model.config.num_heads
        # sequence-parallel runs keep only this rank's heads (realtime_video_b200/parallel.py)
        num_heads = getattr(self.generator.model, "kv_cache_heads", self.generator.model.config.num_heads)
        shape = [batch_size, kv_cache_size, num_heads, head_dim]
        if self.kv_cache1 and list(self.kv_cache1[0]["k"].shape) == shape:
            for c in self.kv_cache1:
                if self.zero_kv_on_reset:
                    c["k"].zero_()
                    c["v"].zero_()
                c["global_end_index"] = 0
                c["local_end_index"] = 0
        else:
            self.kv_cache1 = [{
                "k": torch.zeros(shape, dtype=dtype, device=device),
                "v": torch.zeros(shape, dtype=dtype, device=device),
                "global_end_index": 0, "local_end_index": 0} for _ in range(self.num_transformer_blocks)]
            self.k_shape = self.v_shape = shape

    def _initialize_crossattn_cache(self, batch_size, dtype, device):
        num_heads = self.generator.model.config.num_heads
        dim = self.generator.model.config.dim
        shape = [batch_size, 512, num_heads, dim // num_heads]
        if getattr(self, "crossattn_cache", None) and list(self.crossattn_cache[0]["k"].shape) == shape:
            for c in self.crossattn_cache:
                c["is_init"] = False
        else:
            self.crossattn_cache = [{
                "k": torch.zeros(shape, dtype=dtype, device=device),
                "v": torch.zeros(shape, dtype=dtype, device=device),
                "is_init": False} for _ in range(self.num_transformer_blocks)]
