"""BASELINE configs[3]: 1280x720, kv_cache_num_frames=5 (Lq 10800, Lkv up to 28800), 14B dims.
Runs a few blocks through the session driver and prints per-block GPU time / fps."""
import sys
import time

import torch

sys.path.insert(0, ".")
from realtime_video_b200 import factory, ops  # noqa: E402
import harness  # noqa: E402
from harness import GenerateParams, GenerationSession  # noqa: E402

blocks = int(sys.argv[1]) if len(sys.argv) > 1 else 4
w = factory.synthetic_transformer("14B")
vae = factory.synthetic_vae_decoder()
enc = factory.synthetic_vae_encoder()
models = harness.build_models(w, vae_decoder=vae, vae_encoder=enc)
pe = factory.synthetic_prompt_embeds()
sess = GenerationSession(GenerateParams(width=1280, height=720, kv_cache_num_frames=5, num_blocks=blocks),
                         models, prompt_embeds=pe)
print("frame_seq_length", models.pipeline.frame_seq_length, "kv rows", models.pipeline.kv_cache1[0]["k"].shape[1])
for b in range(blocks):
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    px = sess.generate_block()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e)
    print(f"block {b}: {tuple(px.shape)} {ms:.1f} ms -> {px.shape[1] / ms * 1e3:.2f} fps finite={bool(torch.isfinite(px).all())} "
          f"mem {torch.cuda.max_memory_allocated() / 1e9:.1f} GB", flush=True)
