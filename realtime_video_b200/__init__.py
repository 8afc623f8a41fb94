"""realtime_video_b200 — B200-native (sm_100a) Self-Forcing causal-inference hot path of
krea-ai/realtime-video: causal Wan 2.1 DiT block stack with KV cache + causal 3D VAE decoder,
as hand-written CUDA behind a C ABI (include/krea_b200.h), driven through the reference's own
Python call surface (realtime_video_b200/dropin)."""
__all__ = ["ops", "dit", "session", "factory"]
