"""Multi-GPU single-stream mode on real GPUs (needs >= 2 visible devices; skipped on the one-GPU test box, covered on the
CPU by tests/test_parallel.py with gloo): ``tools/check_sp.py`` under torchrun — the golden 2-layer model on one GPU vs
the same calls sequence-parallel over 2 ranks, once with the kernels writing the rows<->heads exchange into the peers'
buffers over NVLink (``kr_qkv_norm_rope_p2p`` / ``kr_comm_scatter_rows``, KV caches in the symmetric arena) and once
with the NCCL all-to-all baseline: flows and the head-sharded K cache must be bit-identical."""
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_p2p_and_nccl_exchange_are_bit_identical_to_one_gpu():
    import socket
    with socket.socket() as sock:                # a free port: the box may run other rendezvous
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(ROOT / "tools" / "check_sp.py")],
                       capture_output=True, text=True, timeout=300, cwd=str(ROOT))
    out = r.stdout + r.stderr
    assert r.returncode == 0 and out.count("SP CHECK PASS") == 2, out[-3000:]
    assert "bit-identical=False" not in out
