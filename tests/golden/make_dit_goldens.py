"""Generate DiT golden vectors by running the UNMODIFIED reference modules (CPU, build container).

    python tests/golden/make_dit_goldens.py          # writes tests/golden/dit_*.npz

torch %(torch)s; reference krea-ai/realtime-video @ /root/reference.  Synthetic seeded weights
(the reference ships no checkpoint and no test vectors: "parity unpinned" by its own tests,
SURVEY.md §0.6) — these fixtures are what pins our oracle and kernels to the reference code.

Cases (all B=1, dim 256 = 2 heads x 128, ffn 512, 2 layers, text_dim 128):
  cache     : CausalWanModel._forward_inference cache branch: block 0 at two timesteps (the
              second call overwrites the same KV slot), then block 1 attending 6 cached frames.
  recompute : block-causal "flex" branch over 5 context frames (3 frames per block), followed
              by a cache-branch call behind it — the server's recompute_kv_cache sequence
              (release_server.py:588-633).
  evict     : local_attn_size=4, sink_size=1: three blocks through a 4-frame rolling cache.
  wrapper   : WanDiffusionWrapper.forward (flow -> x0 in fp64) + FlowMatchScheduler tables.
  full1560  : UNPATCHED reference at the only self-consistent size (60x104 latent = 1560 tokens
              per frame), 1 layer: block 0 then block 1.
Small-grid cases run the reference with its literal 1560 replaced by 96 (= 8x12 patches), tagged
"patched-constant"; see ref_shim.install(frame_seqlen=...).
Each case is stored for an fp32 run (SDPA fallbacks patched to keep fp32) and a bf16 run (the
reference's own bf16 SDPA fallback).
"""
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
import ref_shim  # noqa: E402

DIMS = dict(dim=256, ffn_dim=512, num_heads=2, num_layers=2, text_dim=128)
OUT = {}


def put(name, t):
    t = t.detach().cpu()
    if t.dtype == torch.bfloat16:
        OUT[name + "@bf16"] = t.contiguous().view(torch.int16).numpy()
    else:
        OUT[name] = t.contiguous().numpy()


def build_model(ns, dtype, **extra):
    torch.manual_seed(0)
    kw = dict(DIMS)
    kw.update(extra)
    m = ns.causal_model.CausalWanModel(**kw)
    torch.manual_seed(1)
    with torch.no_grad():
        # reference init zeroes head.head.weight and all Linear biases (causal_model.py:1151-1173):
        # redraw them so every term of the forward pass is exercised
        m.head.head.weight.normal_(std=0.02)
        for n_, p_ in m.named_parameters():
            if n_.endswith(".bias"):
                p_.normal_(std=0.02)
            if "norm_q.weight" in n_ or "norm_k.weight" in n_ or n_.endswith("norm3.weight"):
                p_.add_(torch.randn_like(p_) * 0.1)
    m = m.to(torch.bfloat16)          # fixture weights are bf16 values
    return m.to(dtype).eval()


def caches(ns, model, size, dtype):
    n, d = model.num_heads, model.dim // model.num_heads
    kv = [{"k": torch.zeros(1, size, n, d, dtype=dtype), "v": torch.zeros(1, size, n, d, dtype=dtype),
           "global_end_index": 0, "local_end_index": 0} for _ in model.blocks]
    ca = [{"k": torch.zeros(1, 512, n, d, dtype=dtype), "v": torch.zeros(1, 512, n, d, dtype=dtype),
           "is_init": False} for _ in model.blocks]
    return kv, ca


@torch.no_grad()
def run_small(tag, dtype):
    FS = 96
    ns = ref_shim.install(frame_seqlen=FS)
    if dtype == torch.float32:
        ref_shim.patch_fp32_sdpa(ns)
    ref_shim.patch_flex_dense(ns)
    model = build_model(ns, dtype)
    if tag == "fp32":
        for k, v in model.state_dict().items():
            put("w/" + k, v.to(torch.bfloat16))
    g = torch.Generator().manual_seed(42)
    ctx = torch.randn(20, 128, generator=g).to(torch.bfloat16)
    xs = [torch.randn(16, 3, 16, 24, generator=g).to(torch.bfloat16) for _ in range(8)]
    x5 = torch.randn(16, 5, 16, 24, generator=g).to(torch.bfloat16)
    if tag == "fp32":
        put("in/ctx", ctx)
        put("in/x5f", x5)
        for i, x in enumerate(xs):
            put(f"in/x{i}", x)

    def fwd(m, x, t, kv, ca, start):
        tt = torch.full((1, x.shape[1]), float(t))
        return m(x[None].to(dtype), t=tt, context=[ctx.to(dtype)], seq_len=32760, kv_cache=kv,
                 crossattn_cache=ca, current_start=start)[0]

    # ---- cache ----
    kv, ca = caches(ns, model, 6 * FS, dtype)
    put(f"{tag}/cache/flow1", fwd(model, xs[0], 1000.0, kv, ca, 0))
    put(f"{tag}/cache/flow2", fwd(model, xs[1], 750.0, kv, ca, 0))
    put(f"{tag}/cache/flow3", fwd(model, xs[2], 1000.0, kv, ca, 3 * FS))
    put(f"{tag}/cache/k0", kv[0]["k"][0])
    put(f"{tag}/cache/v1", kv[1]["v"][0])
    put(f"{tag}/cache/idx", torch.tensor([kv[0]["global_end_index"], kv[0]["local_end_index"]]))

    # ---- recompute (flex branch with the dense mask), then a cache-branch call ----
    kv, ca = caches(ns, model, 8 * FS, dtype)
    mask = ns.causal_model.get_sdpa_mask("cpu", num_frames=5, frame_seqlen=FS, num_frame_per_block=3)
    model.block_mask = mask
    put(f"{tag}/recompute/flow_ctx", fwd(model, x5, 0.0, kv, ca, 5 * FS))
    model.block_mask = None
    put(f"{tag}/recompute/flow_new", fwd(model, xs[3], 1000.0, kv, ca, 5 * FS))
    put(f"{tag}/recompute/k0", kv[0]["k"][0])
    put(f"{tag}/recompute/idx", torch.tensor([kv[0]["global_end_index"], kv[0]["local_end_index"]]))

    # ---- evict ----
    ev = build_model(ns, dtype, local_attn_size=4, sink_size=1)
    kv, ca = caches(ns, ev, 4 * FS, dtype)
    put(f"{tag}/evict/flow1", fwd(ev, xs[4], 1000.0, kv, ca, 0))
    put(f"{tag}/evict/flow2", fwd(ev, xs[5], 1000.0, kv, ca, 3 * FS))
    put(f"{tag}/evict/flow2b", fwd(ev, xs[6], 500.0, kv, ca, 3 * FS))
    put(f"{tag}/evict/flow3", fwd(ev, xs[7], 1000.0, kv, ca, 6 * FS))
    put(f"{tag}/evict/k0", kv[0]["k"][0])
    put(f"{tag}/evict/idx", torch.tensor([kv[0]["global_end_index"], kv[0]["local_end_index"]]))

    # ---- wrapper: flow -> x0 and the scheduler ----
    W = ns.wan_wrapper.WanDiffusionWrapper
    w = W.__new__(W)
    torch.nn.Module.__init__(w)
    w.model = model
    w.uniform_timestep = False
    w.scheduler = ns.scheduler.FlowMatchScheduler(shift=5.0, sigma_min=0.0, extra_one_step=True)
    w.scheduler.set_timesteps(1000, training=True)
    w.seq_len = 32760
    kv, ca = caches(ns, model, 6 * FS, dtype)
    lat = xs[0].permute(1, 0, 2, 3)[None].to(dtype)              # [1, F, C, H, W]
    ts = torch.ones(1, 3, dtype=torch.int64) * 750
    flow, x0 = w(noisy_image_or_video=lat, conditional_dict={"prompt_embeds": ctx[None].to(dtype)},
                 timestep=ts, kv_cache=kv, crossattn_cache=ca, current_start=0)
    put(f"{tag}/wrapper/flow", flow[0])
    put(f"{tag}/wrapper/x0", x0[0])
    if tag == "fp32":
        put("sched/sigmas", w.scheduler.sigmas)
        put("sched/timesteps", w.scheduler.timesteps)
        noise = xs[1].permute(1, 0, 2, 3).to(torch.bfloat16)
        clean = xs[2].permute(1, 0, 2, 3).to(torch.bfloat16)
        put("sched/add_noise_750", w.scheduler.add_noise(clean, noise, torch.full((3,), 750, dtype=torch.long)))
        import importlib
        zp = torch.cat((w.scheduler.timesteps, torch.tensor([0], dtype=torch.float32)))
        for steps in (4, 5):
            lst = torch.linspace(1000.0, 0, steps, dtype=torch.float32).to(torch.long)
            put(f"sched/steps{steps}", zp[1000 - lst])


@torch.no_grad()
def run_full(tag, dtype):
    ns = ref_shim.install(frame_seqlen=None)           # unpatched reference
    if dtype == torch.float32:
        ref_shim.patch_fp32_sdpa(ns)
    # 1-layer model carrying the 2-layer fixture's weights (embeddings, blocks.0, head)
    base = build_model(ns, dtype)
    model = build_model(ns, dtype, num_layers=1)
    model.load_state_dict(base.state_dict(), strict=False)
    g = torch.Generator().manual_seed(7)
    ctx = torch.randn(20, 128, generator=g).to(torch.bfloat16)
    xa = torch.randn(16, 3, 60, 104, generator=g).to(torch.bfloat16)
    xb = torch.randn(16, 3, 60, 104, generator=g).to(torch.bfloat16)
    if tag == "fp32":
        put("full/in/ctx", ctx)
        put("full/in/xa", xa)
        put("full/in/xb", xb)
    kv, ca = caches(ns, model, 6 * 1560, dtype)

    def fwd(x, t, start):
        tt = torch.full((1, 3), float(t))
        return model(x[None].to(dtype), t=tt, context=[ctx.to(dtype)], seq_len=32760, kv_cache=kv,
                     crossattn_cache=ca, current_start=start)[0]

    f1 = fwd(xa, 1000.0, 0)
    f2 = fwd(xb, 500.0, 4680)
    # keep the fixture small: every other row/column
    put(f"{tag}/full/flow1_sub", f1[:, :, ::2, ::2].float().to(torch.float16 if dtype != torch.float32 else torch.float32))
    put(f"{tag}/full/flow2_sub", f2[:, :, ::2, ::2].float().to(torch.float16 if dtype != torch.float32 else torch.float32))
    put(f"{tag}/full/idx", torch.tensor([kv[0]["global_end_index"], kv[0]["local_end_index"]]))


if __name__ == "__main__":
    torch.set_num_threads(8)
    for tag, dt in (("fp32", torch.float32), ("bf16", torch.bfloat16)):
        run_small(tag, dt)
    np.savez_compressed(HERE / "dit_small.npz", **OUT)
    print("dit_small.npz:", len(OUT), "arrays", sum(v.nbytes for v in OUT.values()) / 1e6, "MB raw")
    OUT.clear()
    for tag, dt in (("fp32", torch.float32), ("bf16", torch.bfloat16)):
        run_full(tag, dt)
    np.savez_compressed(HERE / "dit_full1560.npz", **OUT)
    print("dit_full1560.npz:", len(OUT), "arrays", sum(v.nbytes for v in OUT.values()) / 1e6, "MB raw")
