"""Drop-in for ``demo_utils/vae.py``: the single-latent-frame decoder on the B200 kernels.
``VAETRTWrapper`` (TensorRT int8 engine runner, vae.py:317-390) is out of scope
(``use_trt: false`` in both configs)."""
from realtime_video_b200.dropin.demo_utils.constant import ALL_INPUTS_NAMES, ZERO_VAE_CACHE  # noqa: F401
from realtime_video_b200.vae import (AttentionBlock, CausalConv3d, RMS_norm, Resample,  # noqa: F401
                                     ResidualBlock, Upsample, VAEDecoder3d, VAEDecoderWrapperSingle)
