"""Test / bench infrastructure: stand-ins for the reference's two callers of the hot path
(``GenerationSession`` in release_server.py and ``CausalInferencePipeline`` in pipeline/causal_inference.py) for
boxes without the reference checkout.  Nothing in ``realtime_video_b200/`` imports this package; the product is
driven by the reference's own callers (realtime_video_b200/dropin)."""
import types

from .classic_loop import PipelineState
from .server_loop import GenerateParams, GenerationSession, Models, get_denoising_schedule

__all__ = ["PipelineState", "GenerateParams", "GenerationSession", "Models", "get_denoising_schedule",
           "pipeline_args", "build_models"]


def pipeline_args(denoising_step_list=(1000, 750, 500, 250), num_frame_per_block: int = 3):
    """configs/default_config.yaml + self_forcing_server_14b.yaml values the pipeline reads."""
    return types.SimpleNamespace(denoising_step_list=list(denoising_step_list), warp_denoising_step=True,
                                 num_frame_per_block=num_frame_per_block, independent_first_frame=False,
                                 context_noise=0, model_kwargs={})


def build_models(transformer, vae_decoder=None, text_encoder=None, device="cuda", vae_encoder=None) -> Models:
    """release_server.py:227-313 ``load_pipeline`` / ``load_all`` over already-built models."""
    pipe = PipelineState(pipeline_args(), device=device, generator=transformer,
                         text_encoder=text_encoder if text_encoder is not None else object(),
                         vae=vae_decoder if vae_decoder is not None else object())
    return Models(text_encoder, transformer, pipe, vae_encoder, vae_decoder)
