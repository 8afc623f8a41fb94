"""The drop-in directory exposes the reference's import names and call surface
(SURVEY.md §8b): what release_server.py / sample.py import must resolve to our classes when
``realtime_video_b200/dropin`` precedes the reference checkout on sys.path."""
import inspect
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent

SNIPPET = r"""
import sys
sys.path[:0] = [r"%(root)s", r"%(root)s/realtime_video_b200/dropin"]
from pipeline import CausalInferencePipeline                                    # release_server.py:233
from utils.wan_wrapper import WanDiffusionWrapper, WanTextEncoder, WanVAEWrapper  # release_server.py:28
from utils.scheduler import FlowMatchScheduler                                  # release_server.py:556
from demo_utils.vae_block3 import VAEDecoderWrapper, VAEEncoderWrapper          # release_server.py:195, :189
from demo_utils.vae import VAEDecoderWrapperSingle, ZERO_VAE_CACHE, ALL_INPUTS_NAMES
from wan.modules.causal_model import CausalWanModel
import realtime_video_b200.dit as dit, realtime_video_b200.vae as vae
assert CausalWanModel is dit.CausalWanModel and VAEDecoderWrapper is vae.VAEDecoderWrapper
assert VAEEncoderWrapper is vae.VAEEncoderWrapper
import realtime_video_b200.t5 as t5
assert isinstance(WanTextEncoder(model_config=dict(vocab=8, dim=64, dim_attn=64, dim_ffn=64, num_heads=1, num_layers=1, num_buckets=32), device='cpu').text_encoder, t5.T5Encoder)
print("OK")
"""


def test_reference_import_names_resolve_to_the_b200_classes():
    r = subprocess.run([sys.executable, "-c", SNIPPET % {"root": str(ROOT)}], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stderr[-2000:]


def test_call_signatures_match_the_reference():
    from realtime_video_b200.dropin.pipeline.causal_inference import CausalInferencePipeline
    from realtime_video_b200.dropin.utils.wan_wrapper import WanDiffusionWrapper
    from realtime_video_b200.vae import VAEDecoderWrapper, VAEDecoderWrapperSingle, VAEEncoderWrapper
    # utils/wan_wrapper.py:230-242
    fwd = list(inspect.signature(WanDiffusionWrapper.forward).parameters)
    assert fwd[:7] == ["self", "noisy_image_or_video", "conditional_dict", "timestep", "kv_cache",
                       "crossattn_cache", "current_start"]
    assert "cache_start" in fwd
    # utils/wan_wrapper.py:122-131
    init = inspect.signature(WanDiffusionWrapper.__init__).parameters
    for k in ("model_name", "timestep_shift", "is_causal", "local_attn_size", "sink_size", "meta_init"):
        assert k in init
    # pipeline/causal_inference.py:10-17, :48-56
    assert list(inspect.signature(CausalInferencePipeline.__init__).parameters)[:6] == \
        ["self", "args", "device", "generator", "text_encoder", "vae"]
    inf = inspect.signature(CausalInferencePipeline.inference).parameters
    for k in ("noise", "text_prompts", "initial_latent", "return_latents", "profile", "low_memory"):
        assert k in inf
    for name in ("_initialize_kv_cache", "_initialize_crossattn_cache"):
        assert list(inspect.signature(getattr(CausalInferencePipeline, name)).parameters) == \
            ["self", "batch_size", "dtype", "device"]
    # demo_utils/vae_block3.py:195-199, :141-146 ; demo_utils/vae.py:167-172
    assert list(inspect.signature(VAEDecoderWrapper.forward).parameters) == ["self", "z", "feat_cache"]
    assert list(inspect.signature(VAEEncoderWrapper.forward).parameters) == ["self", "z", "feat_cache", "stream"]
    assert list(inspect.signature(VAEDecoderWrapperSingle.forward).parameters) == \
        ["self", "z", "is_first_frame", "feat_cache"]


def test_session_context_frame_selection_matches_reference_rule():
    """release_server.py:563-576: first frame + last kv-1 frames while (block_idx-1)*3 < kv, afterwards
    the re-encoded oldest pixel frame replaces the first latent frame."""
    import torch
    from realtime_video_b200.session import GenerateParams, GenerationSession
    s = GenerationSession.__new__(GenerationSession)
    s.params = GenerateParams(kv_cache_num_frames=3, keep_first_frame=True)
    s.models = type("M", (), {"pipeline": type("P", (), {"num_frame_per_block": 3})(), "vae_encoder": None})()
    s.decode_enabled = True
    s.all_latents = torch.arange(12.0).view(1, 12, 1, 1, 1)
    for block_idx, start, expect in [(1, 3, [0, 1, 2]), (2, 6, [0, 4, 5]), (3, 9, [0, 7, 8])]:
        s.block_idx, s.current_start_frame = block_idx, start
        assert s.get_clean_context_frames().flatten().tolist() == expect
