"""Drop-in for ``demo_utils/constant.py``: cache placeholders.  The B200 decoder keeps its
feature cache in channels-last buffers of its own; an all-None list means "new stream"."""
ENCODER_ZERO_VAE_CACHE = [None] * 55
DECODER_ZERO_VAE_CACHE = [None] * 55
ZERO_VAE_CACHE = DECODER_ZERO_VAE_CACHE
ALL_INPUTS_NAMES = ["z", "is_first_frame"] + [f"cache_{i}" for i in range(32)]
