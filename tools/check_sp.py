"""torchrun --nproc-per-node N tools/check_sp.py : sequence-parallel DiT == single-GPU DiT.

Every rank first runs the golden 2-layer model alone (sp=None), then the same calls with SequenceParallel over all
ranks — once with the kernels doing the exchange over NVLink peer memory ("p2p": kr_qkv_norm_rope_p2p /
kr_comm_scatter_rows, KV caches in the symmetric arena) and once with the NCCL all-to-all baseline — and compares
flows and (its head slice of) the KV cache.  Same arithmetic per row / head, so the results are bit-identical unless
a GEMM shape switches to the stream-K kernel (then: fp32 summation order, rel-L2 <= 1e-3)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from realtime_video_b200.parallel import SequenceParallel  # noqa: E402
from tests import test_dit_gpu as T  # noqa: E402
from tests.golden_io import load_npz, rel_l2  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
g = load_npz("dit_small.npz")
FS = 96
heads = 2 if world <= 2 else 4


def build():
    if heads == 2:
        return T.build(g)
    from realtime_video_b200.dit import CausalWanModel
    torch.manual_seed(5)
    m = CausalWanModel(dim=512, ffn_dim=1024, num_heads=4, num_layers=2, text_dim=128)
    with torch.no_grad():
        m.head.head.weight.normal_(std=0.02)
    return m.to(device="cuda", dtype=torch.bfloat16).eval()


def caches(m, size, nheads):
    kv = [{"k": torch.zeros(1, size, nheads, 128, dtype=torch.bfloat16, device="cuda"),
           "v": torch.zeros(1, size, nheads, 128, dtype=torch.bfloat16, device="cuda"),
           "global_end_index": 0, "local_end_index": 0} for _ in m.blocks]
    ca = [{"k": torch.zeros(1, 512, m.num_heads, 128, dtype=torch.bfloat16, device="cuda"),
           "v": torch.zeros(1, 512, m.num_heads, 128, dtype=torch.bfloat16, device="cuda"),
           "is_init": False} for _ in m.blocks]
    return kv, ca


def run(m, kv, ca):
    outs = []
    outs.append(T.fwd(m, g, "in/x0", 1000, kv, ca, 0))
    outs.append(T.fwd(m, g, "in/x1", 750, kv, ca, 0))
    outs.append(T.fwd(m, g, "in/x2", 1000, kv, ca, 3 * FS))
    m.block_mask = m._prepare_blockwise_causal_attn_mask("cuda", num_frames=3, frame_seqlen=FS,
                                                         num_frame_per_block=3, local_attn_size=-1)
    for c in kv:
        c["global_end_index"] = c["local_end_index"] = 0
    outs.append(T.fwd(m, g, "in/x3", 0, kv, ca, 3 * FS))
    m.block_mask = None
    outs.append(T.fwd(m, g, "in/x4", 1000, kv, ca, 3 * FS))
    return outs


m = build()
kv1, ca1 = caches(m, 6 * FS, m.num_heads)
ref = run(m, kv1, ca1)
ok = True
for exchange in ("p2p", "nccl"):
    m.sp = SequenceParallel(exchange=exchange)
    hl = m.kv_cache_heads
    if exchange == "p2p":
        kvs = m.sp.alloc_kv_cache(len(m.blocks), (1, 6 * FS, hl, 128), torch.bfloat16, torch.device("cuda", local))
        if kvs is None:       # symmetric memory unavailable: every rank has fallen back to the collective exchange
            print(f"rank {rank}: p2p setup failed ({m.sp.fallback_reason}); checking the fallback exchange instead")
            kv2, ca2 = caches(m, 6 * FS, hl)
        else:
            _, ca2 = caches(m, 6 * FS, hl)
            kv2 = [{"k": k, "v": v, "global_end_index": 0, "local_end_index": 0} for k, v in kvs]
    else:
        kv2, ca2 = caches(m, 6 * FS, hl)
    got = run(m, kv2, ca2)
    for i, (a, b) in enumerate(zip(got, ref)):
        same = torch.equal(a, b)
        ok &= same or rel_l2(a, b) < 1e-3
        print(f"[{exchange} rank {rank}/{world}] call {i}: bit-identical={same} rel_l2={rel_l2(a, b):.2e}", flush=True)
    k_ref = kv1[1]["k"][0][:, rank * hl:(rank + 1) * hl]
    same = torch.equal(kv2[1]["k"][0], k_ref)
    r = rel_l2(kv2[1]["k"][0], k_ref)
    print(f"[{exchange} rank {rank}] layer-1 K cache head slice bit-identical={same} rel_l2={r:.2e}", flush=True)
    ok &= same or r < 1e-3
    torch.cuda.synchronize()
    dist.barrier()
    m.sp = None
dist.barrier()
dist.destroy_process_group()
print(f"[rank {rank}] SP CHECK {'PASS' if ok else 'FAIL'}", flush=True)
sys.exit(0 if ok else 1)
