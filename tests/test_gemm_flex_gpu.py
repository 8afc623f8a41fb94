"""The wave-fitted single-CTA GEMM (gemm_flex_kernel: runtime tile width chosen so the tile count fills whole waves
of SMs, last column tile narrower) on the row shards of the multi-GPU mode, every epilogue, against an fp32 matmul
of the same bf16 operands: rel-L2 <= 2e-3."""
import pytest
import torch

from tests.golden_io import rel_l2

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M", [585, 1170, 2340])
@pytest.mark.parametrize("N,K,epi", [(15360, 5120, "split"), (5120, 5120, "gate_res"), (13824, 5120, "gelu"),
                                     (5120, 13824, "gate_res"), (5120, 5120, "bias"), (4096, 4096, "mul")])
def test_flex_width_gemm(M, N, K, epi):
    from realtime_video_b200 import _lib, ops
    code = {"bias": 0, "gelu": 1, "gate_res": 2, "split": 0, "mul": 5}[epi]
    if _lib.load().kr_gemm_kernel_id_ws(code, M, N, K, 0) != 4:
        pytest.skip("the planner keeps a fixed-width kernel for this shape")
    ops_sk, ops.stream_k = ops.stream_k, False
    try:
        torch.manual_seed(M + N + K)
        a = torch.randn(M, K, device="cuda").bfloat16()
        w = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
        b = torch.randn(N, device="cuda").bfloat16()
        ref = a.float() @ w.float().t() + b.float()
        if epi == "bias":
            out, want = ops.gemm(a, w, b), ref
        elif epi == "gelu":
            out = ops.gemm(a, w, b, epilogue=ops.EPI_BIAS_GELU)
            want = torch.nn.functional.gelu(ref.bfloat16().float(), approximate="tanh")
        elif epi == "mul":
            g = torch.randn(M, N, device="cuda").bfloat16()
            out = ops.gemm(a, w, b, epilogue=ops.EPI_MUL, residual=g)
            want = ref.bfloat16().float() * g.float()
        elif epi == "gate_res":
            x = torch.ones(M, N, device="cuda").bfloat16()
            gate = torch.randn(3, N, device="cuda").bfloat16()
            rpg = (M + 2) // 3
            rows = torch.arange(M, device="cuda") // rpg
            want = x.float() + (ref.bfloat16().float() * gate.float()[rows]).bfloat16().float()
            out = ops.gemm(a, w, b, epilogue=ops.EPI_BIAS_GATE_RES, residual=x, gate=gate, rows_per_gate=rpg, out=x)
        else:
            out = torch.empty(M, N - N // 3, device="cuda", dtype=torch.bfloat16)
            v = torch.zeros(M + 64, N // 3, device="cuda", dtype=torch.bfloat16)
            ops.gemm(a, w, b, out=out, out2=v[32:32 + M], n_split=N - N // 3)
            assert rel_l2(v[32:32 + M].float(), ref[:, N - N // 3:]) < 2e-3
            assert float(v[:32].abs().max()) == 0 and float(v[32 + M:].abs().max()) == 0
            want = ref[:, :N - N // 3]
        torch.cuda.synchronize()
        assert rel_l2(out.float(), want) < 2e-3
    finally:
        ops.stream_k = ops_sk
