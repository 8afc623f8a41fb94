"""First contact of the whole DiT path with the GPU: small-model parity numbers + 14B timing."""
import sys
import time

import torch

sys.path.insert(0, ".")
from realtime_video_b200 import factory, ops  # noqa: E402
import harness  # noqa: E402
from harness import GenerateParams, GenerationSession  # noqa: E402


def parity():
    from tests import test_dit_gpu as T
    from tests.golden_io import load_npz, rel_l2
    g = load_npz("dit_small.npz")
    m = T.build(g)
    kv, ca = T.caches(m, 6 * 96)
    for (x, t, s, name) in [("in/x0", 1000, 0, "cache/flow1"), ("in/x1", 750, 0, "cache/flow2"),
                            ("in/x2", 1000, 288, "cache/flow3")]:
        out = T.fwd(m, g, x, t, kv, ca, s)
        print(name, "vs fp32", rel_l2(out, g["fp32/" + name]), "vs bf16", rel_l2(out, g["bf16/" + name]),
              "| ref bf16 vs fp32", rel_l2(g["bf16/" + name], g["fp32/" + name]), flush=True)


def speed(layers):
    t0 = time.time()
    w = factory.synthetic_transformer("14B", num_layers=layers)
    torch.cuda.synchronize()
    print(f"built {layers}-layer 14B-dims model in {time.time() - t0:.1f}s, mem {torch.cuda.memory_allocated() / 1e9:.1f} GB", flush=True)
    models = harness.build_models(w)
    pe = factory.synthetic_prompt_embeds()
    sess = GenerationSession(GenerateParams(num_blocks=4), models, prompt_embeds=pe, decode=False)
    for b in range(4):
        torch.cuda.synchronize()
        s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
        t1 = time.time()
        n0 = ops.launch_count
        s.record()
        out = sess.generate_block()
        e.record()
        torch.cuda.synchronize()
        print(f"block {b}: gpu {s.elapsed_time(e):.1f} ms, wall {1e3 * (time.time() - t1):.1f} ms, "
              f"launches {ops.launch_count - n0}, finite={bool(torch.isfinite(out.float()).all())} "
              f"absmax={out.float().abs().max().item():.3f}", flush=True)


if __name__ == "__main__":
    which = sys.argv[1]
    if which == "parity":
        parity()
    else:
        speed(int(sys.argv[2]))
