"""Host logic of realtime_video_b200/dit.py executed on the CPU: the kernels behind ``ops`` are replaced by the
fp32 stand-ins of tests/cpu_ops_emulation.py, everything else (cache-slot arithmetic, rolling eviction with sink
tokens, fused-QKV split output, recompute branch with the block mask and padded keys, cross-attention cache,
modulation indexing, unpatchify, flow -> x0) is the product code.  Compared with the goldens produced by the
reference's own modules in fp32 (tests/golden/dit_small.npz): both sides are fp32, so rel-L2 <= 1e-4."""
import pytest
import torch

from tests import cpu_ops_emulation as emu
from tests.golden_io import load_npz, rel_l2, weights

FS = 96
TOL = 1e-4


@pytest.fixture(scope="module")
def g():
    return load_npz("dit_small.npz")


@pytest.fixture(autouse=True)
def cpu_ops(monkeypatch):
    import realtime_video_b200.dit as dit
    import realtime_video_b200.wan_wrapper as ww
    import realtime_video_b200.vae as vae
    for mod in (dit, ww, vae):
        monkeypatch.setattr(mod, "ops", emu)


def build(g, num_layers=2, **kw):
    from realtime_video_b200.dit import CausalWanModel
    m = CausalWanModel(dim=256, ffn_dim=512, num_heads=2, num_layers=num_layers, text_dim=128, **kw)
    missing = m.load_state_dict(weights(g, torch.float32), strict=False)
    assert not missing.missing_keys, missing
    return m.float().eval()


def caches(m, size):
    n, d = m.num_heads, m.dim // m.num_heads
    kv = [{"k": torch.zeros(1, size, n, d), "v": torch.zeros(1, size, n, d),
           "global_end_index": 0, "local_end_index": 0} for _ in m.blocks]
    ca = [{"k": torch.zeros(1, 512, n, d), "v": torch.zeros(1, 512, n, d), "is_init": False} for _ in m.blocks]
    return kv, ca


def fwd(m, g, xname, t, kv, ca, start):
    x = g[xname].float()
    tt = torch.full((1, x.shape[1]), float(t))
    with torch.no_grad():
        return m(x[None], t=tt, context=g["in/ctx"].float()[None], seq_len=32760, kv_cache=kv, crossattn_cache=ca,
                 current_start=start)[0]


def check(out, g, name):
    r = rel_l2(out, g[f"fp32/{name}"])
    assert r < TOL, f"{name}: rel_l2={r:.3e}"


@pytest.mark.parametrize("fused", [False, True])
def test_cache_branch(g, fused):
    m = build(g)
    if fused:
        for blk in m.blocks:
            blk.self_attn.fuse_projections()          # exercises the split (V -> cache slot) GEMM output
    kv, ca = caches(m, 6 * FS)
    check(fwd(m, g, "in/x0", 1000, kv, ca, 0), g, "cache/flow1")
    check(fwd(m, g, "in/x1", 750, kv, ca, 0), g, "cache/flow2")
    check(fwd(m, g, "in/x2", 1000, kv, ca, 3 * FS), g, "cache/flow3")
    check(kv[0]["k"][0], g, "cache/k0")
    check(kv[1]["v"][0], g, "cache/v1")
    assert [kv[0]["global_end_index"], kv[0]["local_end_index"]] == g["bf16/cache/idx"].tolist()
    assert all(c["is_init"] for c in ca)


def test_model_without_cross_attn_norm_keeps_the_residual_stream(g):
    """cross_attn_norm=False: norm3 is nn.Identity (causal_model.py:424-426) and the cross-attention reads the residual
    stream itself; the FFN pre-norm must NOT be written over it (regression: it used to be, through an aliased scratch
    buffer).  No Wan 2.1 checkpoint uses this configuration, so the check is against the oracle with the same flag."""
    from oracle import dit_oracle as O
    m = build(g, cross_attn_norm=False)
    w = weights(g, torch.float32)
    orc = O.DiTOracle(O.DiTConfig(dim=256, ffn_dim=512, num_heads=2, num_layers=2, text_dim=128,
                                  frame_seqlen_const=FS, cross_attn_norm=False), w)
    kv, ca = caches(m, 6 * FS)
    kv_o = O.new_kv_cache(orc.cfg, 6 * FS, torch.float32)
    ca_o = O.new_crossattn_cache(orc.cfg, torch.float32)
    for xname, t, start in (("in/x0", 1000, 0), ("in/x2", 1000, 3 * FS)):
        got = fwd(m, g, xname, t, kv, ca, start)
        ref = orc.forward_inference(g[xname].float(), torch.full((3,), float(t)), g["in/ctx"].float(), kv_o, ca_o, start)
        assert rel_l2(got, ref) < TOL


def test_recompute_branch_with_block_mask_and_padded_keys(g):
    m = build(g)
    kv, ca = caches(m, 8 * FS)
    m.block_mask = m._prepare_blockwise_causal_attn_mask("cpu", num_frames=5, frame_seqlen=FS,
                                                         num_frame_per_block=3, local_attn_size=-1)
    check(fwd(m, g, "in/x5f", 0, kv, ca, 5 * FS), g, "recompute/flow_ctx")
    m.block_mask = None
    check(fwd(m, g, "in/x3", 1000, kv, ca, 5 * FS), g, "recompute/flow_new")
    check(kv[0]["k"][0], g, "recompute/k0")
    assert [kv[0]["global_end_index"], kv[0]["local_end_index"]] == g["bf16/recompute/idx"].tolist()


def test_rolling_eviction_with_sink(g):
    m = build(g, local_attn_size=4, sink_size=1)
    kv, ca = caches(m, 4 * FS)
    check(fwd(m, g, "in/x4", 1000, kv, ca, 0), g, "evict/flow1")
    check(fwd(m, g, "in/x5", 1000, kv, ca, 3 * FS), g, "evict/flow2")
    check(fwd(m, g, "in/x6", 500, kv, ca, 3 * FS), g, "evict/flow2b")
    check(fwd(m, g, "in/x7", 1000, kv, ca, 6 * FS), g, "evict/flow3")
    check(kv[0]["k"][0], g, "evict/k0")
    assert [kv[0]["global_end_index"], kv[0]["local_end_index"]] == g["bf16/evict/idx"].tolist()


def test_cache_overflow_is_an_error(g):
    m = build(g)
    kv, ca = caches(m, 3 * FS)
    fwd(m, g, "in/x0", 1000, kv, ca, 0)
    with pytest.raises(RuntimeError, match="KV cache overflow"):
        fwd(m, g, "in/x2", 1000, kv, ca, 3 * FS)


def test_wrapper_flow_and_x0(g):
    from realtime_video_b200.wan_wrapper import WanDiffusionWrapper
    w = WanDiffusionWrapper(model_name="synthetic", timestep_shift=5.0, is_causal=True,
                            model_config=dict(dim=256, ffn_dim=512, num_heads=2, num_layers=2, text_dim=128))
    w.model.load_state_dict(weights(g, torch.float32), strict=False)
    w = w.float().eval()
    kv, ca = caches(w.model, 6 * FS)
    lat = g["in/x0"].float().permute(1, 0, 2, 3)[None].contiguous()
    ts = torch.ones(1, 3, dtype=torch.int64) * 750
    with torch.no_grad():
        flow, x0 = w(noisy_image_or_video=lat, conditional_dict={"prompt_embeds": g["in/ctx"].float()[None]},
                     timestep=ts, kv_cache=kv, crossattn_cache=ca, current_start=0)
    assert flow.shape == lat.shape and x0.shape == lat.shape
    check(flow[0], g, "wrapper/flow")
    check(x0[0], g, "wrapper/x0")


def test_wrapper_batch_of_two_equals_two_single_calls(g):
    """B > 1 (the classic pipeline accepts a batch of noise, pipeline/causal_inference.py:48-50): every sample uses its
    own slice of the [B, rows, heads, 128] KV cache and its own prompt; results and cache contents must equal two
    batch-1 calls, and the shared python indices advance once."""
    from realtime_video_b200.wan_wrapper import WanDiffusionWrapper
    w = WanDiffusionWrapper(model_name="synthetic", timestep_shift=5.0, is_causal=True,
                            model_config=dict(dim=256, ffn_dim=512, num_heads=2, num_layers=2, text_dim=128))
    w.model.load_state_dict(weights(g, torch.float32), strict=False)
    w = w.float().eval()
    n, d = 2, 128
    lat = torch.stack([g["in/x0"].float().permute(1, 0, 2, 3), g["in/x1"].float().permute(1, 0, 2, 3)]).contiguous()
    ctx = torch.stack([g["in/ctx"].float(), g["in/ctx"].float().flip(0)])
    ts = torch.ones(2, 3, dtype=torch.int64) * 750

    def kv_cache(B):
        return [{"k": torch.zeros(B, 6 * FS, n, d), "v": torch.zeros(B, 6 * FS, n, d), "global_end_index": 0,
                 "local_end_index": 0} for _ in w.model.blocks]

    def ca_cache(B):
        return [{"k": torch.zeros(B, 512, n, d), "v": torch.zeros(B, 512, n, d), "is_init": False} for _ in w.model.blocks]

    kv2 = kv_cache(2)
    with torch.no_grad():
        flow2, x02 = w(noisy_image_or_video=lat, conditional_dict={"prompt_embeds": ctx}, timestep=ts, kv_cache=kv2,
                       crossattn_cache=ca_cache(2), current_start=0)
    assert flow2.shape == lat.shape
    for b in range(2):
        kv1 = kv_cache(1)
        with torch.no_grad():
            flow1, x01 = w(noisy_image_or_video=lat[b:b + 1], conditional_dict={"prompt_embeds": ctx[b:b + 1]},
                           timestep=ts[b:b + 1], kv_cache=kv1, crossattn_cache=ca_cache(1), current_start=0)
        assert torch.equal(flow2[b], flow1[0]) and torch.equal(x02[b], x01[0])
        for c2, c1 in zip(kv2, kv1):
            assert torch.equal(c2["k"][b], c1["k"][0]) and torch.equal(c2["v"][b], c1["v"][0])
            assert (c2["global_end_index"], c2["local_end_index"]) == (c1["global_end_index"], c1["local_end_index"])


def test_classic_inference_loop_host_logic_vs_reference():
    """CausalInferencePipeline.inference (2 blocks x (4 denoise + 1 context) passes with re-noising) on the CPU
    against the UNMODIFIED reference pipeline (tests/golden/pipeline_small.npz): cache allocation, per-block
    start offsets, timestep handling, the context pass and WanVAEWrapper.decode_to_pixel are product code.
    Latents: fp32 on both sides, rel-L2 <= 1e-3; video: the VAE engine stores fp16/bf16 activations, mean |d|
    <= 5e-3 on pixels in [0, 1]."""
    import types

    from harness import PipelineState as CausalInferencePipeline
    from realtime_video_b200.wan_wrapper import WanDiffusionWrapper, WanVAEWrapper
    from realtime_video_b200.factory import synthetic_vae_params
    g = load_npz("pipeline_small.npz")
    gd = load_npz("dit_small.npz")
    gen = WanDiffusionWrapper(model_name="synthetic", timestep_shift=5.0, is_causal=True,
                              model_config=dict(dim=256, ffn_dim=512, num_heads=2, num_layers=2, text_dim=128))
    gen.model.load_state_dict(weights(gd, torch.float32), strict=False)
    gen = gen.float().eval()

    vae = WanVAEWrapper(load_pretrained=False)
    vae.model.load_state_dict(synthetic_vae_params(seed=0), strict=False)
    vae = vae.half().eval()

    ctx = g["ctx"].float()
    args = types.SimpleNamespace(denoising_step_list=[1000, 750, 500, 250], warp_denoising_step=True,
                                 num_frame_per_block=3, independent_first_frame=False, context_noise=0,
                                 model_kwargs={})
    pipe = CausalInferencePipeline(args, "cpu", generator=gen,
                                   text_encoder=lambda text_prompts: {"prompt_embeds": ctx}, vae=vae)
    pipe.frame_seq_length = 96          # (H/16)*(W/16); the reference's literal 1560 is 832x480
    assert torch.allclose(pipe.denoising_step_list.float(), g["steps"].float())
    it = iter([g[f"draw{i}"] for i in range(6)])
    real = torch.randn_like
    torch.randn_like = lambda t, **kw: next(it).to(device=t.device, dtype=t.dtype)
    try:
        video, latents = pipe.inference(g["noise"].float(), ["x"], return_latents=True)
    finally:
        torch.randn_like = real
    r = rel_l2(latents, g["latents"])
    assert r < 1e-3, f"latents rel_l2={r:.3e}"
    assert video.shape == (1, 21, 3, 128, 192) and float(video.min()) >= 0.0 and float(video.max()) <= 1.0
    mad = (video[..., ::2, ::2] - g["video_sub"]).abs().mean().item()
    assert mad < 5e-3, f"video mean|d|={mad:.3e}"
    assert pipe.kv_cache1[0]["k"].shape[1] == 32760        # the reference allocates its literal (:289)
