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


@pytest.mark.parametrize("world,L,heads", [(2, 24, 4), (4, 40, 4), (8, 64, 8)])
def test_rows_heads_all_to_all_gloo(world, L, heads):
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, L, heads, ret), nprocs=world, join=True)
    assert [ret.get(r) for r in range(world)] == ["ok"] * world


def _sp_model_worker(rank, world, port, ret):
    """The sequence-parallel branch of dit.py (row shards, head-sharded KV cache, two all-to-alls per layer,
    row offsets of the modulation / RoPE / gate lookups) on `world` gloo ranks, kernels replaced by the fp32
    stand-ins: every rank must reproduce the reference's fp32 goldens of the single-GPU schedule."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import realtime_video_b200.dit as dit
        from realtime_video_b200.parallel import SequenceParallel
        from tests import cpu_ops_emulation as emu
        from tests.golden_io import load_npz, rel_l2, weights
        dit.ops = emu
        g = load_npz("dit_small.npz")
        FS = 96
        m = dit.CausalWanModel(dim=256, ffn_dim=512, num_heads=2, num_layers=2, text_dim=128)
        m.load_state_dict(weights(g, torch.float32), strict=False)
        m = m.float().eval()
        for blk in m.blocks:
            blk.self_attn.fuse_projections()
        m.sp = SequenceParallel()
        n, d = m.kv_cache_heads, m.dim // m.num_heads
        assert n == 2 // world
        kv = [{"k": torch.zeros(1, 6 * FS, n, d), "v": torch.zeros(1, 6 * FS, n, d),
               "global_end_index": 0, "local_end_index": 0} for _ in m.blocks]
        ca = [{"k": torch.zeros(1, 512, 2, d), "v": torch.zeros(1, 512, 2, d), "is_init": False} for _ in m.blocks]

        def fwd(xname, t, start):
            x = g[xname].float()
            with torch.no_grad():
                return m(x[None], t=torch.full((1, x.shape[1]), float(t)), context=g["in/ctx"].float()[None],
                         seq_len=32760, kv_cache=kv, crossattn_cache=ca, current_start=start)[0]

        for xname, t, start, name in (("in/x0", 1000, 0, "cache/flow1"), ("in/x1", 750, 0, "cache/flow2"),
                                      ("in/x2", 1000, 3 * FS, "cache/flow3")):
            r = rel_l2(fwd(xname, t, start), g[f"fp32/{name}"])
            assert r < 1e-4, (name, r)
        # my head block of the reference's K cache
        k_ref = g["fp32/cache/k0"].reshape(-1, 2, d)[:, rank * n:(rank + 1) * n]
        assert rel_l2(kv[0]["k"][0], k_ref) < 1e-4
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_sequence_parallel_dit_schedule_gloo():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_sp_model_worker, args=(world, port, ret), nprocs=world, join=True)
    assert [ret.get(r) for r in range(world)] == ["ok"] * world


def _fallback_worker(rank, world, port, ret):
    """p2p setup failing on ONE rank only must move EVERY rank to the collective exchange (a split decision would
    hang the first layer): rank 0 gets a working fake symmetric allocation, rank 1 raises."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import types
        import warnings

        import torch.distributed._symmetric_memory as symm_mem
        from realtime_video_b200.parallel import SequenceParallel

        def fake_empty(n, dtype=None, device=None):
            if rank == 1:
                raise RuntimeError("no peer access")
            return torch.empty(n, dtype=dtype)

        def fake_rendezvous(t, group_name):
            return types.SimpleNamespace(buffer_ptrs=[t.data_ptr(), 0], barrier=lambda channel=0: None)

        symm_mem.empty, symm_mem.rendezvous = fake_empty, fake_rendezvous
        sp = SequenceParallel(exchange="p2p")
        assert sp.p2p
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            kv = sp.alloc_kv_cache(2, (1, 48, 1, 128), torch.float32, "cpu")
        assert kv is None and sp.exchange == "nccl" and not sp.p2p and sp.fallback_reason
        assert any("symmetric-memory setup failed" in str(x.message) for x in w)
        # the collective exchange works afterwards
        x = torch.arange(24 * 256, dtype=torch.float32).view(24, 256)
        r0, n = sp.rows(24)
        assert torch.equal(sp.rows_to_heads(x[r0:r0 + n]), x[:, rank * 128:(rank + 1) * 128])
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_p2p_setup_failure_on_one_rank_falls_back_everywhere_gloo():
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_fallback_worker, args=(2, port, ret), nprocs=2, join=True)
    assert [ret.get(r) for r in range(2)] == ["ok"] * 2


def _pp_model_worker(rank, world, port, ret):
    """BASELINE configs[2]: the DiT layers sharded over the ranks (LayerPipeline), residual stream sent rank -> rank + 1,
    head on the last rank, result broadcast.  Golden 2-layer model on 2 gloo ranks (one layer each), kernels replaced
    by the fp32 stand-ins: every rank must return the reference's fp32 goldens, and each rank may only have touched
    the caches of ITS layer."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import realtime_video_b200.dit as dit
        from realtime_video_b200.parallel import LayerPipeline
        from tests import cpu_ops_emulation as emu
        from tests.golden_io import load_npz, rel_l2, weights
        dit.ops = emu
        g = load_npz("dit_small.npz")
        FS = 96
        m = dit.CausalWanModel(dim=256, ffn_dim=512, num_heads=2, num_layers=2, text_dim=128)
        m.load_state_dict(weights(g, torch.float32), strict=False)
        m = m.float().eval()
        m.pp = LayerPipeline()
        assert m.pp.layers(2) == (rank, rank + 1) and m.pp.layers(5) == ((0, 3) if rank == 0 else (3, 5))
        n, d = 2, m.dim // m.num_heads
        kv = [{"k": torch.zeros(1, 6 * FS, n, d), "v": torch.zeros(1, 6 * FS, n, d),
               "global_end_index": 0, "local_end_index": 0} for _ in m.blocks]
        ca = [{"k": torch.zeros(1, 512, n, d), "v": torch.zeros(1, 512, n, d), "is_init": False} for _ in m.blocks]

        def fwd(xname, t, start):
            x = g[xname].float()
            with torch.no_grad():
                return m(x[None], t=torch.full((1, x.shape[1]), float(t)), context=g["in/ctx"].float()[None],
                         seq_len=32760, kv_cache=kv, crossattn_cache=ca, current_start=start)[0]

        for xname, t, start, name in (("in/x0", 1000, 0, "cache/flow1"), ("in/x1", 750, 0, "cache/flow2"),
                                      ("in/x2", 1000, 3 * FS, "cache/flow3")):
            r = rel_l2(fwd(xname, t, start), g[f"fp32/{name}"])
            assert r < 1e-4, (name, r)
        mine, other = rank, 1 - rank
        assert ca[mine]["is_init"] and not ca[other]["is_init"]
        assert kv[mine]["local_end_index"] == 6 * FS and kv[other]["local_end_index"] == 0
        assert float(kv[other]["k"].abs().max()) == 0.0 and float(kv[mine]["k"].abs().max()) > 0.0
        if rank == 0:
            assert rel_l2(kv[0]["k"][0], g["fp32/cache/k0"]) < 1e-4
        else:
            assert rel_l2(kv[1]["v"][0], g["fp32/cache/v1"]) < 1e-4
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_layer_pipeline_dit_schedule_gloo():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_pp_model_worker, args=(world, port, ret), nprocs=world, join=True)
    assert [ret.get(r) for r in range(world)] == ["ok"] * world


def _pp_session_worker(rank, world, port, ret):
    """The server block loop (harness.GenerationSession) with the DiT layer-pipelined over 2 gloo ranks: rank 0 decodes
    and re-encodes the first context frame for everybody (broadcast), both ranks must carry the same latents as a
    single process, block by block, through the sliding window."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import realtime_video_b200.dit as dit
        import realtime_video_b200.vae as vae
        import realtime_video_b200.wan_wrapper as ww
        from realtime_video_b200.parallel import LayerPipeline
        from tests import cpu_ops_emulation as emu
        for mod in (dit, ww, vae):
            mod.ops = emu
        from tests.test_session_host_cpu import make

        def run(pipelined):
            sess, _ = make(keep_first_frame=False, decode=(not pipelined) or rank == 0, seed=11, blocks=3)
            model = sess.models.pipeline.generator.model
            model.pp = LayerPipeline() if pipelined else None
            if pipelined:           # the session was built before the mode was set: recompute what depends on it
                sess.decode_enabled = True
            outs = []
            for _ in range(3):          # block 2 is the first one behind the sliding window (first-frame re-encode)
                out = sess.generate_block()
                outs.append((out, sess.all_latents.clone()))
            return outs

        single = run(False)
        piped = run(True)
        for b, ((o1, lat1), (o2, lat2)) in enumerate(zip(single, piped)):
            assert torch.equal(lat1, lat2), f"block {b}: latents differ on rank {rank}"
            if rank == 0:
                assert torch.equal(o1, o2), f"block {b}: pixels differ"
            else:
                assert o2.shape[-3] == 16            # non-decoding ranks hand back the latents of the block
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_layer_pipeline_server_loop_gloo():
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_pp_session_worker, args=(2, port, ret), nprocs=2, join=True)
    assert [ret.get(r) for r in range(2)] == ["ok"] * 2
