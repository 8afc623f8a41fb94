"""VAE decoder on the GPU: timing at 832x480 (latent 60x104), 3 latent frames per call."""
import sys
import time

import torch

sys.path.insert(0, ".")
from realtime_video_b200.factory import synthetic_vae_params  # noqa: E402
from realtime_video_b200 import ops  # noqa: E402
from realtime_video_b200.vae import VAEDecoderWrapper  # noqa: E402

m = VAEDecoderWrapper()
m.load_state_dict(synthetic_vae_params(seed=0), strict=False)
m = m.to(device="cuda", dtype=torch.float16).eval()
cache = [None] * 55
g = torch.Generator(device="cuda").manual_seed(0)
with torch.no_grad():
    for call in range(5):
        z = torch.randn(1, 3, 16, 60, 104, device="cuda", generator=g).half()
        torch.cuda.synchronize()
        s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
        n0 = ops.launch_count
        t0 = time.time()
        s.record()
        px, cache = m(z, *cache)
        e.record()
        torch.cuda.synchronize()
        print(f"call {call}: {tuple(px.shape)} gpu {s.elapsed_time(e):.1f} ms wall {1e3 * (time.time() - t0):.1f} ms "
              f"launches {ops.launch_count - n0} finite={bool(torch.isfinite(px).all())} "
              f"clamped={float((px.abs() >= 1).float().mean()):.3f} mem {torch.cuda.max_memory_allocated() / 1e9:.1f} GB",
              flush=True)
