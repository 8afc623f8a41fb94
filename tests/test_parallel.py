"""Sequence-parallel plumbing on CPU: world_size-2 (and 4) gloo process groups.

Checks that the row<->head all-to-all layout of realtime_video_b200/parallel.py reproduces, shard
by shard, what a single process computes: attention over head-sharded q/k/v equals the matching
head block of full attention, and the round trip rows->heads->rows is the identity."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, L, heads, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.dit_oracle import attention
        from realtime_video_b200.parallel import SequenceParallel
        sp = SequenceParallel()
        g = torch.Generator().manual_seed(0)               # same tensors on every rank
        q = torch.randn(L, heads * 128, generator=g)
        k = torch.randn(L, heads * 128, generator=g)
        v = torch.randn(L, heads * 128, generator=g)
        r0, n = sp.rows(L)
        hl = sp.local_heads(heads)
        assert (r0, n) == (rank * L // world, L // world) and hl == heads // world
        # rows -> heads: every rank ends up with ALL rows of ITS heads, in global row order
        kcache = torch.zeros(L + 7, hl * 128)               # receive straight into a "cache slot"
        qh = sp.rows_to_heads(q[r0:r0 + n])
        sp.rows_to_heads(k[r0:r0 + n], out=kcache[3:3 + L])
        vh = sp.rows_to_heads(v[r0:r0 + n])
        cols = slice(rank * hl * 128, (rank + 1) * hl * 128)
        assert torch.equal(qh, q[:, cols]) and torch.equal(kcache[3:3 + L], k[:, cols]) and torch.equal(vh, v[:, cols])
        # attention on my heads == my head block of full attention
        full = attention(q.view(L, heads, 128), k.view(L, heads, 128), v.view(L, heads, 128)).reshape(L, -1)
        mine = attention(qh.view(L, hl, 128), kcache[3:3 + L].view(L, hl, 128), vh.view(L, hl, 128)).reshape(L, -1)
        assert torch.allclose(mine, full[:, cols], atol=1e-6)
        # heads -> rows: back to my rows with all heads
        back = sp.heads_to_rows(mine)
        assert torch.allclose(back, full[r0:r0 + n], atol=1e-6)
        assert torch.equal(sp.heads_to_rows(qh), q[r0:r0 + n])
        # gather_rows
        assert torch.equal(sp.gather_rows(q[r0:r0 + n, :64].contiguous()), q[:, :64])
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,L,heads", [(2, 24, 4), (4, 40, 4)])
def test_rows_heads_all_to_all_gloo(world, L, heads):
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, L, heads, ret), nprocs=world, join=True)
    assert [ret.get(r) for r in range(world)] == ["ok"] * world
