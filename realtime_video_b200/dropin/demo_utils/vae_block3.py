"""Drop-in for ``demo_utils/vae_block3.py``: the server's VAE decoder (vae_block3.py:177-230) and streaming VAE
encoder (vae_block3.py:116-175) on the B200 kernels."""
from realtime_video_b200.vae import (AttentionBlock, CausalConv3d, Encoder3d, RMS_norm, Resample,  # noqa: F401
                                     ResidualBlock, Upsample, VAEDecoder3d, VAEDecoderWrapper, VAEEncoderWrapper)
