"""Pin oracle/dit_oracle.py against golden vectors produced by the reference's own modules
(tests/golden/make_dit_goldens.py).  CPU only."""
import pytest
import torch

from oracle import dit_oracle as O
from tests.golden_io import load_npz, rel_l2, weights

FS = 96
TOL = {"fp32": 2e-5, "bf16": 1.5e-2}   # bf16: same op order, different matmul summation order


@pytest.fixture(scope="module")
def g():
    return load_npz("dit_small.npz")


def make(g, tag, **kw):
    dt = torch.float32 if tag == "fp32" else torch.bfloat16
    cfg = O.DiTConfig(dim=256, ffn_dim=512, num_heads=2, num_layers=2, text_dim=128,
                      frame_seqlen_const=FS, **kw)
    return O.DiTOracle(cfg, weights(g, dt)), cfg, dt


def run(m, g, dt, xname, t, kv, ca, start, mask_args=None):
    x = g[xname].to(dt)
    tt = torch.full((x.shape[1],), float(t))
    return m.forward_inference(x, tt, g["in/ctx"].to(dt), kv, ca, start, mask_args)


@pytest.mark.parametrize("tag", ["fp32", "bf16"])
def test_cache_branch(g, tag):
    m, cfg, dt = make(g, tag)
    kv, ca = O.new_kv_cache(cfg, 6 * FS, dt), O.new_crossattn_cache(cfg, dt)
    assert rel_l2(run(m, g, dt, "in/x0", 1000, kv, ca, 0), g[f"{tag}/cache/flow1"]) < TOL[tag]
    assert rel_l2(run(m, g, dt, "in/x1", 750, kv, ca, 0), g[f"{tag}/cache/flow2"]) < TOL[tag]
    assert rel_l2(run(m, g, dt, "in/x2", 1000, kv, ca, 3 * FS), g[f"{tag}/cache/flow3"]) < TOL[tag]
    assert rel_l2(kv[0]["k"][0], g[f"{tag}/cache/k0"]) < TOL[tag]
    assert rel_l2(kv[1]["v"][0], g[f"{tag}/cache/v1"]) < TOL[tag]
    assert [kv[0]["global_end_index"], kv[0]["local_end_index"]] == g[f"{tag}/cache/idx"].tolist()


@pytest.mark.parametrize("tag", ["fp32", "bf16"])
def test_recompute_branch(g, tag):
    m, cfg, dt = make(g, tag)
    kv, ca = O.new_kv_cache(cfg, 8 * FS, dt), O.new_crossattn_cache(cfg, dt)
    out = run(m, g, dt, "in/x5f", 0, kv, ca, 5 * FS, mask_args={"block_len": 3 * FS})
    assert rel_l2(out, g[f"{tag}/recompute/flow_ctx"]) < TOL[tag]
    out = run(m, g, dt, "in/x3", 1000, kv, ca, 5 * FS)
    assert rel_l2(out, g[f"{tag}/recompute/flow_new"]) < TOL[tag]
    assert rel_l2(kv[0]["k"][0], g[f"{tag}/recompute/k0"]) < TOL[tag]
    assert [kv[0]["global_end_index"], kv[0]["local_end_index"]] == g[f"{tag}/recompute/idx"].tolist()


@pytest.mark.parametrize("tag", ["fp32", "bf16"])
def test_eviction_branch(g, tag):
    m, cfg, dt = make(g, tag, local_attn_size=4, sink_size=1)
    kv, ca = O.new_kv_cache(cfg, 4 * FS, dt), O.new_crossattn_cache(cfg, dt)
    assert rel_l2(run(m, g, dt, "in/x4", 1000, kv, ca, 0), g[f"{tag}/evict/flow1"]) < TOL[tag]
    assert rel_l2(run(m, g, dt, "in/x5", 1000, kv, ca, 3 * FS), g[f"{tag}/evict/flow2"]) < TOL[tag]
    assert rel_l2(run(m, g, dt, "in/x6", 500, kv, ca, 3 * FS), g[f"{tag}/evict/flow2b"]) < TOL[tag]
    assert rel_l2(run(m, g, dt, "in/x7", 1000, kv, ca, 6 * FS), g[f"{tag}/evict/flow3"]) < TOL[tag]
    assert rel_l2(kv[0]["k"][0], g[f"{tag}/evict/k0"]) < TOL[tag]
    assert [kv[0]["global_end_index"], kv[0]["local_end_index"]] == g[f"{tag}/evict/idx"].tolist()


@pytest.mark.parametrize("tag", ["fp32", "bf16"])
def test_wrapper_flow_to_x0(g, tag):
    m, cfg, dt = make(g, tag)
    kv, ca = O.new_kv_cache(cfg, 6 * FS, dt), O.new_crossattn_cache(cfg, dt)
    flow = run(m, g, dt, "in/x0", 750, kv, ca, 0).permute(1, 0, 2, 3)       # [F, C, H, W]
    assert rel_l2(flow, g[f"{tag}/wrapper/flow"]) < TOL[tag]
    s = O.FlowMatchSchedulerOracle(shift=5.0)
    xt = g["in/x0"].to(dt).permute(1, 0, 2, 3)
    # conversion checked on the reference's own flow so the comparison is exact
    x0 = O.flow_to_x0(g[f"{tag}/wrapper/flow"].to(dt), xt, torch.full((3,), 750, dtype=torch.int64), s)
    assert torch.equal(x0, g[f"{tag}/wrapper/x0"].to(dt))


@pytest.mark.parametrize("tag", ["fp32", "bf16"])
def test_full_1560_unpatched(tag):
    """The reference with its hard-coded 1560 tokens/frame at the one self-consistent size."""
    g = load_npz("dit_full1560.npz")
    small = load_npz("dit_small.npz")
    dt = torch.float32 if tag == "fp32" else torch.bfloat16
    cfg = O.DiTConfig(dim=256, ffn_dim=512, num_heads=2, num_layers=1, text_dim=128)
    m = O.DiTOracle(cfg, weights(small, dt))
    kv, ca = O.new_kv_cache(cfg, 6 * 1560, dt), O.new_crossattn_cache(cfg, dt)
    ctx = g["full/in/ctx"].to(dt)
    f1 = m.forward_inference(g["full/in/xa"].to(dt), torch.full((3,), 1000.0), ctx, kv, ca, 0)
    f2 = m.forward_inference(g["full/in/xb"].to(dt), torch.full((3,), 500.0), ctx, kv, ca, 4680)
    tol = TOL[tag] if tag == "fp32" else 2e-2
    assert rel_l2(f1[:, :, ::2, ::2], g[f"{tag}/full/flow1_sub"]) < tol
    assert rel_l2(f2[:, :, ::2, ::2], g[f"{tag}/full/flow2_sub"]) < tol
    assert [kv[0]["global_end_index"], kv[0]["local_end_index"]] == g[f"{tag}/full/idx"].tolist()


def test_scheduler_tables(g):
    s = O.FlowMatchSchedulerOracle(shift=5.0)
    assert torch.allclose(s.sigmas, g["sched/sigmas"], rtol=0, atol=1e-7)
    assert torch.allclose(s.timesteps, g["sched/timesteps"], rtol=0, atol=1e-4)
    for steps in (4, 5):
        assert torch.equal(O.denoising_schedule(s, 1.0, steps), g[f"sched/steps{steps}"])
    clean = g["in/x2"].permute(1, 0, 2, 3)
    noise = g["in/x1"].permute(1, 0, 2, 3)
    got = s.add_noise(clean, noise, torch.full((3,), 750, dtype=torch.long))
    assert torch.equal(got, g["sched/add_noise_750"])
