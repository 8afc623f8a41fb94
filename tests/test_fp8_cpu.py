"""FP8 row on the CPU: the oracle's quantisation invariants, ``fp8.quantize_`` on the DiT module
tree, the torchao import names served by the drop-in finder, and the block schedule with FP8 linears (stand-in
kernels) against the bf16 schedule."""
import pytest
import torch

from oracle import fp8_oracle as F8
from tests import cpu_ops_emulation as emu
from tests.golden_io import load_npz, rel_l2, weights


def test_oracle_quantisation_invariants():
    torch.manual_seed(0)
    x = torch.randn(64, 256) * 3
    q, s = F8.quantize_per_tensor(x)
    assert q.dtype == torch.float8_e4m3fn and float(q.float().abs().max()) == 448.0      # amax maps onto the e4m3 max
    back = q.float() * s
    assert rel_l2(back, x) < 4e-2                                                        # 3 mantissa bits
    w = torch.randn(128, 256) * 0.02
    y = F8.linear_fp8(x.bfloat16(), w.bfloat16(), torch.zeros(128).bfloat16())
    assert y.dtype == torch.bfloat16 and rel_l2(y.float(), x @ w.t()) < 6e-2


def test_quantize_attaches_fp8_weights_to_the_block_linears_only():
    from realtime_video_b200 import fp8
    from realtime_video_b200.wan_wrapper import WanDiffusionWrapper
    w = WanDiffusionWrapper(model_name="synthetic", is_causal=True,
                            model_config=dict(dim=256, ffn_dim=512, num_heads=2, num_layers=2, text_dim=128))
    for blk in w.model.blocks:
        blk.self_attn.fuse_projections()
    keys = set(w.state_dict().keys())
    fp8.quantize_(w, fp8.Float8DynamicActivationFloat8WeightConfig(granularity=fp8.PerTensor()))
    blk = w.model.blocks[0]
    for lin in (blk.self_attn.to_qkv, blk.self_attn.o, blk.cross_attn.q, blk.cross_attn.k, blk.ffn[0], blk.ffn[2]):
        wq, sw = lin._kr_fp8
        assert wq.dtype == torch.uint8 and wq.shape == lin.weight.shape and sw > 0
        ref_q, ref_s = F8.quantize_per_tensor(lin.weight.detach())
        assert torch.equal(wq.view(torch.float8_e4m3fn).float(), ref_q.float()) and abs(sw - float(ref_s)) < 1e-12
    assert not hasattr(w.model.head.head, "_kr_fp8") and not hasattr(w.model.time_embedding[0], "_kr_fp8")
    assert set(w.state_dict().keys()) == keys                         # checkpoint surface unchanged
    fp8.dequantize_(w)
    assert not hasattr(blk.ffn[0], "_kr_fp8")


def test_server_fp8_import_lines_resolve_through_the_dropin():
    import subprocess
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    code = ("import realtime_video_b200.dropin as d; d.install()\n"
            "from torchao.quantization.quant_api import quantize_, Float8DynamicActivationFloat8WeightConfig, PerTensor\n"
            "import realtime_video_b200.fp8 as f\n"
            "assert quantize_ is f.quantize_ and PerTensor is f.PerTensor; print('OK')\n")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True,
                       env={"PYTHONPATH": str(root), "PATH": "/usr/bin:/bin"})
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stderr[-2000:]


def test_block_schedule_with_fp8_linears_tracks_the_bf16_schedule(monkeypatch):
    """Host logic: with ``quantize_`` applied the schedule calls ``linear_fp8`` for every block linear (incl. the split
    QKV output into the V cache slot); results stay within the FP8 tier's tolerance of the bf16 schedule."""
    import realtime_video_b200.dit as dit
    from realtime_video_b200 import fp8
    monkeypatch.setattr(dit, "ops", emu)
    g = load_npz("dit_small.npz")
    m = dit.CausalWanModel(dim=256, ffn_dim=512, num_heads=2, num_layers=2, text_dim=128)
    m.load_state_dict(weights(g, torch.float32), strict=False)
    m = m.float().eval()
    for blk in m.blocks:
        blk.self_attn.fuse_projections()

    def run():
        kv = [dict(k=torch.zeros(1, 6 * 96, 2, 128), v=torch.zeros(1, 6 * 96, 2, 128), global_end_index=0,
                   local_end_index=0) for _ in range(2)]
        ca = [dict(k=torch.zeros(1, 512, 2, 128), v=torch.zeros(1, 512, 2, 128), is_init=False) for _ in range(2)]
        x = g["in/x0"].float()
        with torch.no_grad():
            return m(x[None], t=torch.full((1, 3), 750.0), context=g["in/ctx"].float()[None], seq_len=32760,
                     kv_cache=kv, crossattn_cache=ca, current_start=0)[0], kv
    ref, _ = run()
    calls = {"n": 0}
    real = emu.linear_fp8

    def counted(*a, **k):
        calls["n"] += 1
        return real(*a, **k)
    monkeypatch.setattr(emu, "linear_fp8", counted)
    fp8.quantize_(m)
    got, kv = run()
    assert calls["n"] == 2 * 8                      # per block: to_qkv, o, cross q/k/v/o, ffn.0, ffn.2
    assert kv[0]["local_end_index"] == 3 * 96 and float(kv[0]["v"].abs().sum()) > 0
    r = rel_l2(got, ref)
    assert r < 0.15, r                               # two blocks of e4m3 linears on a 256-wide toy model


@pytest.mark.parametrize("M,N,K", [(32, 48, 64), (96, 256, 512), (17, 128, 1024)])
def test_oracle_linear_matches_torch_scaled_mm(M, N, K):
    """torchao's Float8 dynamic-activation linear ends in ``torch._scaled_mm(x_q, w_q.t(), scale_a, scale_b, bias,
    out_dtype=bf16)`` — the call the reference's `enable_fp8` path executes (release_server.py:179-182).  torchao itself
    is absent here, but that kernel's semantics are available on the CPU: the oracle's scaled matmul + bias + bf16
    rounding is pinned to it (the per-tensor scale formula stays a restatement of torchao's published algorithm)."""
    if not hasattr(torch, "_scaled_mm"):
        pytest.skip("torch._scaled_mm not available")
    torch.manual_seed(M + N + K)
    x, w, b = torch.randn(M, K) * 2, torch.randn(N, K) * 0.03, torch.randn(N)
    xq, sx = F8.quantize_per_tensor(x)
    wq, sw = F8.quantize_per_tensor(w)
    try:
        want = torch._scaled_mm(xq, wq.t(), scale_a=sx.reshape(1).float(), scale_b=sw.reshape(1).float(),
                                bias=b.bfloat16(), out_dtype=torch.bfloat16)
    except (RuntimeError, NotImplementedError) as ex:
        pytest.skip(f"torch._scaled_mm unsupported on this CPU build: {ex}")
    got = F8.linear_fp8(x, w, b.bfloat16())
    assert got.dtype == want.dtype == torch.bfloat16
    assert rel_l2(got.float(), want.float()) < 2e-3
    assert (got == want).float().mean() > 0.98            # identical up to fp32 summation order before the bf16 rounding
