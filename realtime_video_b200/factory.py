"""Build the hot-path models with seeded synthetic weights (there are no checkpoints on the
build or GPU boxes): the reference's init (xavier / N(0,.02)) with
``head.head.weight`` redrawn N(0,.02) because the reference zero-initialises it
(causal_model.py:1173) and a zero head makes every output identically zero."""
from __future__ import annotations

import types

import torch

from .dropin.pipeline.causal_inference import CausalInferencePipeline
from .dropin.utils.wan_wrapper import KNOWN_CONFIGS, WanDiffusionWrapper
from .session import Models


def synthetic_transformer(size: str = "14B", device="cuda", dtype=torch.bfloat16, seed: int = 0,
                          timestep_shift: float = 5.0, num_layers: int | None = None,
                          **overrides) -> WanDiffusionWrapper:
    cfg = dict(KNOWN_CONFIGS[size]) if size in KNOWN_CONFIGS else {}
    cfg.update(overrides)
    if num_layers is not None:
        cfg["num_layers"] = num_layers
    torch.manual_seed(seed)
    if torch.device(device).type == "cuda":
        torch.cuda.manual_seed(seed)
    w = WanDiffusionWrapper(model_name="synthetic-" + size, timestep_shift=timestep_shift, is_causal=True,
                            model_config=cfg, device=device, dtype=dtype)
    with torch.no_grad():
        w.model.head.head.weight.normal_(std=0.02)
    w.eval().requires_grad_(False)
    for blk in w.model.blocks:          # release_server.py:176-177
        blk.self_attn.fuse_projections()
    return w


def synthetic_prompt_embeds(device="cuda", text_dim: int = 4096, tokens: int = 64, seed: int = 1):
    """randn [1, 512, text_dim] with rows >= tokens zeroed (utils/wan_wrapper.py:52-53)."""
    g = torch.Generator().manual_seed(seed)
    e = torch.randn(1, 512, text_dim, generator=g)
    e[:, tokens:] = 0
    return e.to(device=device, dtype=torch.bfloat16)


def pipeline_args(denoising_step_list=(1000, 750, 500, 250), num_frame_per_block: int = 3):
    """configs/default_config.yaml + self_forcing_server_14b.yaml values the pipeline reads."""
    return types.SimpleNamespace(denoising_step_list=list(denoising_step_list), warp_denoising_step=True,
                                 num_frame_per_block=num_frame_per_block, independent_first_frame=False,
                                 context_noise=0, model_kwargs={})


def build_models(transformer: WanDiffusionWrapper, vae_decoder=None, text_encoder=None, device="cuda") -> Models:
    pipe = CausalInferencePipeline(pipeline_args(), device=device, generator=transformer,
                                   text_encoder=text_encoder if text_encoder is not None else object(),
                                   vae=vae_decoder if vae_decoder is not None else object())
    return Models(text_encoder, transformer, pipe, None, vae_decoder)
