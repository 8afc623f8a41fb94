"""Drop-in for ``demo_utils/vae_block3.py``: the server VAE decoder on the B200 kernels.
``VAEEncoderWrapper`` (streaming encoder, vae_block3.py:116-175) is a "next" row
and is taken from the reference tree when that is importable."""
from realtime_video_b200.dropin.utils.wan_wrapper import _reference_module
from realtime_video_b200.vae import (AttentionBlock, CausalConv3d, RMS_norm, Resample,  # noqa: F401
                                     ResidualBlock, Upsample, VAEDecoder3d, VAEDecoderWrapper)


def __getattr__(name):
    if name in ("VAEEncoderWrapper", "VAEEncoder3d"):
        import os
        mod = _reference_module("demo_utils.vae_block3", os.path.join("demo_utils", "vae_block3.py"))
        if mod is None:
            raise ImportError(f"{name} is outside the B200 hot path (SURVEY.md §8f.1); put the reference "
                              f"checkout on sys.path after realtime_video_b200/dropin")
        return getattr(mod, name)
    raise AttributeError(name)
