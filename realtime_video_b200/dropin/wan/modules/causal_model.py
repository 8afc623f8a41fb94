"""Drop-in for ``wan.modules.causal_model``: re-exports the B200 model classes."""
from realtime_video_b200.dit import (BlockMaskSpec, CausalHead, CausalWanAttentionBlock,  # noqa: F401
                                     CausalWanModel, CausalWanSelfAttention)
