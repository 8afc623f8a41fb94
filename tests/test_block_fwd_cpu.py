"""kr_dit_block_fwd (one C-ABI call per DiT block, realtime_video_b200/csrc/kr_dit_block.cu) issues EXACTLY the launches
of the per-op schedule in realtime_video_b200/dit.py — proven without a GPU:

the real kr_api.cu + kr_dit_block.cu + kr_host.cu are linked against recording stand-ins for the kernels' host
launchers (tests/kr_record_stubs.cu), so every launch becomes a log line with all its arguments.  The same block is
then driven (a) through the per-op Python schedule (14 ctypes calls) and (b) through ONE kr_dit_block_fwd call, and the
two launch sequences are compared call by call: same launcher, same scalars, same external tensors at the same offsets,
and — for temporaries, which live at different addresses in the two runs — the same DATAFLOW (every read resolves to
the same producing launch and offset).  Identical launches on identical data give identical results on the GPU, which
is how the fused entry point inherits the parity of the per-op path (tests/test_zz_block_fwd_gpu.py checks it live).
Covers the cache branch (first / later block), the rolling-window eviction, the block-causal recompute branch and
a model without the affine norm3."""
import ctypes
import shutil
import subprocess
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "realtime_video_b200" / "csrc"
pytestmark = pytest.mark.skipif(shutil.which("nvcc") is None and not Path("/usr/local/cuda/bin/nvcc").exists(),
                                reason="nvcc not available (host-only compile of the API layer)")


@pytest.fixture(scope="module")
def reclib(tmp_path_factory):
    from realtime_video_b200 import _lib
    so = tmp_path_factory.mktemp("krrec") / "libkrea_record.so"
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    subprocess.run([nvcc, "-std=c++17", "-O1", "-shared", "-Xcompiler", "-fPIC", "-cudart", "static",
                    "-Wno-deprecated-gpu-targets", str(CSRC / "kr_host.cu"), str(CSRC / "kr_api.cu"),
                    str(CSRC / "kr_dit_block.cu"), str(ROOT / "tests" / "kr_record_stubs.cu"), "-o", str(so)], check=True)
    lib = ctypes.CDLL(str(so))
    for name, argtypes in _lib.SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = ctypes.c_int, argtypes
    lib.kr_dit_block_workspace_bytes.restype = ctypes.c_size_t
    lib.kr_dit_block_workspace_bytes.argtypes = [ctypes.c_int] * 4
    lib.kr_gemm_workspace_bytes.restype = ctypes.c_size_t
    lib.kr_last_error.restype = ctypes.c_char_p
    lib.kr_record_dump.restype = ctypes.c_char_p
    return lib


@pytest.fixture()
def on_recorder(reclib, monkeypatch):
    from realtime_video_b200 import ops
    monkeypatch.setattr(ops._lib, "load", lambda: reclib)
    monkeypatch.setattr(ops, "_req", lambda t, name, dtype=None: None)
    monkeypatch.setattr(ops, "_stream", lambda: 0x5EED)
    monkeypatch.setattr(ops, "stream_k", False)
    monkeypatch.setattr(ops, "_block_ws", {})
    return reclib


# ----------------------------------------------------------------------------------------------------------------
# log parsing + dataflow normalisation
# ----------------------------------------------------------------------------------------------------------------
# launcher -> (input pointer fields, output fields with (rows, ld, cols) extent in elements)
ROLES = {
    "gemm": (["a", "w", "bias", "residual", "gate", "ws"],
             {"out": lambda c: (c["M"], c["ldc"], c["n_split"] if c["out2"] else c["N"]),
              "out2": lambda c: (c["M"], c["ldc2"], c["N"] - c["n_split"])}),
    "attn": (["q", "k", "v"], {"out": lambda c: (c["Lq"], c["ldo"], c["heads"] * 128)}),
    "ln": (["x", "w", "b", "mod"], {"out": lambda c: (c["rows"], c["ldo"], c["D"])}),
    "qkv_post": (["q", "k", "v", "wq", "wk", "rope"],
                 {"q_out": lambda c: (c["rows"], c["ldqo"], c["D"]), "k_out": lambda c: (c["rows"], c["ldko"], c["D"]),
                  "v_out": lambda c: (c["rows"], c["ldvo"], c["D"])}),
    "rmsnorm": (["x", "w"], {"out": lambda c: (c["rows"], c["ldo"], c["D"])}),
    "add_mod": (["modulation", "e0"], {"out": lambda c: (1, 0, c["frames"] * c["mod_rows"] * c["D"])}),
    "kv_roll": (["cache"], {}),
}


def parse(log: str):
    calls = []
    for line in log.strip().splitlines():
        name, *kv = line.split()
        c = {"fn": name}
        for item in kv:
            k, v = item.split("=", 1)
            if v == "(nil)":
                c[k] = 0
            elif v.startswith("0x"):
                c[k] = int(v, 16)
            else:
                c[k] = float(v) if ("." in v or "e" in v or "inf" in v) else int(v)
        calls.append(c)
    return calls


def normalise(calls, externals):
    """externals: name -> (address, bytes).  Returns the address-free form of the launch sequence."""
    writes = []                    # (start, end, call index, field) of writes to non-external memory, newest last

    def resolve(ptr):
        if not ptr:
            return None
        for name, (base, nbytes) in externals.items():
            if base <= ptr < base + nbytes:
                return ("ext", name, ptr - base)
        for start, end, idx, field in reversed(writes):
            if start <= ptr < end:
                return ("tmp", idx, field, ptr - start)
        raise AssertionError(f"launch reads memory nobody wrote: {hex(ptr)}")

    out = []
    for i, c in enumerate(calls):
        ins, outs = ROLES[c["fn"]]
        norm = {"fn": c["fn"]}
        ptr_fields = set(ins) | set(outs)
        for k, v in c.items():
            if k not in ptr_fields and k != "fn":
                norm[k] = v
        for f in ins:
            norm[f] = resolve(c.get(f, 0))
        pending = []
        for f, extent in outs.items():
            ptr = c.get(f, 0)
            if not ptr:
                norm[f] = None
                continue
            r = None
            for name, (base, nbytes) in externals.items():
                if base <= ptr < base + nbytes:
                    r = ("ext", name, ptr - base)
            if r is None:
                rows, ld, cols = extent(c)
                pending.append((ptr, ptr + 2 * ((rows - 1) * ld + cols), i, f))
                r = ("tmp-out",)
            norm[f] = r
        writes.extend(pending)       # a launch's own outputs are visible to LATER launches only
        out.append(norm)
    return out


# ----------------------------------------------------------------------------------------------------------------
# the model under test
# ----------------------------------------------------------------------------------------------------------------
GH, GW = 4, 8
FS = GH * GW            # 32 tokens per frame


def build(cross_attn_norm=True, local_attn_size=-1):
    from realtime_video_b200.dit import CausalWanModel
    torch.manual_seed(0)
    m = CausalWanModel(dim=256, ffn_dim=512, num_heads=2, num_layers=1, text_dim=128, cross_attn_norm=cross_attn_norm,
                       local_attn_size=local_attn_size)
    m = m.to(torch.bfloat16).eval()
    for blk in m.blocks:
        blk.self_attn.fuse_projections()
    return m


def run_both(lib, m, frames, kv_rows, current_start, mask, g_end=0, l_end=0):
    """Record the per-op schedule and the one-call schedule of block 0 on the same tensors."""
    blk = m.blocks[0]
    L, D = frames * FS, m.dim
    x = torch.randn(L, D).bfloat16()
    e0 = torch.randn(frames, 6, D).bfloat16()
    kv = {"k": torch.zeros(1, kv_rows, 2, 128, dtype=torch.bfloat16), "v": torch.zeros(1, kv_rows, 2, 128, dtype=torch.bfloat16)}
    ca = {"k": torch.randn(1, 512, 2, 128).bfloat16(), "v": torch.randn(1, 512, 2, 128).bfloat16(), "is_init": True}
    rope = m._rope(torch.device("cpu"))
    ext = {"x": x, "e0": e0, "kc": kv["k"], "vc": kv["v"], "ck": ca["k"], "cv": ca["v"], "rope": rope}
    for n, prm in blk.named_parameters():
        ext["p:" + n] = prm.data
    externals = {n: (t.data_ptr(), t.numel() * t.element_size()) for n, t in ext.items()}
    logs, ends = [], []
    for one_call in (False, True):
        m.use_block_fwd = one_call
        kv["global_end_index"], kv["local_end_index"] = g_end, l_end
        m.block_mask = mask
        lib.kr_record_clear()
        with torch.no_grad():
            m._block(blk, x, e0, (frames, GH, GW), None, kv, ca, current_start, mask)
        logs.append(parse(lib.kr_record_dump().decode()))
        ends.append((kv["global_end_index"], kv["local_end_index"]))
    assert ends[0] == ends[1]
    # The eviction memmove (kv_roll on the K and V caches) runs inside _self_attention on the per-op path (after
    # add_modulation and the first LayerNorm) and in front of the C call on the one-call path.  The launches it moves
    # across never touch the caches (asserted here), so the two orders are the same program: compare with the
    # kv_roll launches hoisted to the front.
    cache_ranges = [externals["kc"], externals["vc"]]
    for lg in logs:
        first_roll = next((i for i, c in enumerate(lg) if c["fn"] == "kv_roll"), None)
        if first_roll:
            for c in lg[:first_roll]:
                for v in c.values():
                    assert not any(isinstance(v, int) and base <= v < base + n for base, n in cache_ranges), c
        lg.sort(key=lambda c: c["fn"] != "kv_roll")          # stable: kv_roll first, everything else in order
    return [normalise(lg, externals) for lg in logs], logs


def check_equal(norm, logs, n_launches=14):
    a, b = norm
    assert [c["fn"] for c in a] == [c["fn"] for c in b]
    for i, (ca_, cb_) in enumerate(zip(a, b)):
        assert ca_ == cb_, f"launch {i} ({ca_['fn']}) differs:\nper-op  {ca_}\none-call {cb_}"
    assert len(a) == n_launches
    assert all(c.get("stream", 0x5EED) == 0x5EED for c in logs[1])


def test_cache_branch_first_block(on_recorder):
    m = build()
    norm, logs = run_both(on_recorder, m, frames=3, kv_rows=6 * FS, current_start=0, mask=None)
    check_equal(norm, logs)
    fns = [c["fn"] for c in norm[1]]
    assert fns == ["add_mod", "ln", "gemm", "qkv_post", "attn", "gemm", "ln", "gemm", "rmsnorm", "attn", "gemm", "ln",
                   "gemm", "gemm"]
    # the V third of the fused projection lands in the cache slot, K comes from the RMSNorm+RoPE launch
    assert norm[1][2]["out2"] == ("ext", "vc", 0) and norm[1][3]["k_out"] == ("ext", "kc", 0)


def test_cache_branch_later_block_attends_the_whole_prefix(on_recorder):
    m = build()
    norm, logs = run_both(on_recorder, m, frames=3, kv_rows=6 * FS, current_start=3 * FS, mask=None, g_end=3 * FS,
                          l_end=3 * FS)
    check_equal(norm, logs)
    attn = norm[1][4]
    assert attn["Lkv"] == 6 * FS and attn["k"] == ("ext", "kc", 0) and norm[1][3]["k_out"] == ("ext", "kc", 3 * FS * 256 * 2)
    assert norm[1][3]["start_frame"] == 3


def test_rolling_window_eviction(on_recorder):
    m = build(local_attn_size=4)                       # 4-frame window, cache full -> the new 3 frames evict 3
    for blk in m.blocks:
        blk.self_attn.local_attn_size = 4
    norm, logs = run_both(on_recorder, m, frames=3, kv_rows=4 * FS, current_start=4 * FS, mask=None, g_end=4 * FS,
                          l_end=4 * FS)
    check_equal(norm, logs, n_launches=16)             # two kv_roll launches (K and V) in front
    assert [c["fn"] for c in norm[1][:2]] == ["kv_roll", "kv_roll"] and norm[1][0]["rows"] == FS


def test_recompute_branch_block_causal(on_recorder):
    m = build()
    mask = m._prepare_blockwise_causal_attn_mask("cpu", num_frames=3, frame_seqlen=FS, num_frame_per_block=3,
                                                 local_attn_size=-1)
    norm, logs = run_both(on_recorder, m, frames=3, kv_rows=6 * FS, current_start=3 * FS, mask=mask)
    check_equal(norm, logs)
    attn = norm[1][4]
    assert attn["mask_mode"] == 1 and attn["block_len"] == 3 * FS and attn["pad_keys"] == 128 - 3 * FS and attn["Lkv"] == 3 * FS


def test_model_without_affine_norm3(on_recorder):
    m = build(cross_attn_norm=False)
    norm, logs = run_both(on_recorder, m, frames=3, kv_rows=6 * FS, current_start=0, mask=None)
    check_equal(norm, logs, n_launches=13)             # no LayerNorm in front of the cross-attention


def test_ineligible_blocks_stay_on_the_per_op_path(on_recorder):
    """FP8 weights attached / prompt K/V not projected yet / profiling on: the one-call path must step aside."""
    m = build()
    m.use_block_fwd = True
    blk = m.blocks[0]
    x = torch.zeros(FS, 256).bfloat16()
    assert m._block_fwd_eligible(blk, x, {"is_init": True})
    assert not m._block_fwd_eligible(blk, x, {"is_init": False})
    assert not m._block_fwd_eligible(blk, x, None)
    blk.ffn[0]._kr_fp8 = (None, 1.0)
    assert not m._block_fwd_eligible(blk, x, {"is_init": True})
    del blk.ffn[0]._kr_fp8
    blk.self_attn.fused_projections = False
    assert not m._block_fwd_eligible(blk, x, {"is_init": True})


def test_argument_validation(reclib):
    from realtime_video_b200 import _lib
    p = _lib.KrDitBlockParams()
    assert reclib.kr_dit_block_fwd(ctypes.byref(p), None) == -1 and b"kr_dit_block_fwd" in reclib.kr_last_error()
    assert reclib.kr_dit_block_workspace_bytes(4680, 5120, 13824, 3) >= 4680 * (5 * 5120 + 13824) * 2 + 3 * 6 * 5120 * 2
    assert reclib.kr_dit_block_workspace_bytes(0, 5120, 13824, 3) == 0


# ----------------------------------------------------------------------------------------------------------------
# the multi-GPU exchange at 8 ranks: host-side pointer arithmetic (never run on 8 GPUs by the builder)
# ----------------------------------------------------------------------------------------------------------------
class _FakeSP:
    """SequenceParallel with the p2p exchange, minus torch.distributed: a fake table of per-rank arena base addresses
    stands for the symmetric-memory rendezvous, the barrier is a no-op."""
    exchange, p2p, fallback_reason, group = "p2p", True, None, None

    def __init__(self, rank, world, L, D):
        from realtime_video_b200.parallel import SequenceParallel
        self.rank, self.world = rank, world
        self.rows = lambda n: SequenceParallel.rows(self, n)
        self.local_heads = lambda h: SequenceParallel.local_heads(self, h)
        dh = D // world
        self.arena = torch.zeros(4 * L * D + 4096, dtype=torch.bfloat16)
        self.q_buf = self.arena[:L * dh].view(L, dh)
        self.o_rows = self.arena[L * dh:L * dh + (L // world) * D].view(L // world, D)
        self.o_heads = torch.zeros(L, dh, dtype=torch.bfloat16)
        self.kv_off = L * dh + (L // world) * D
        self.peer_base = [0x10_0000_0000 * (r + 1) for r in range(world)]
        self.barriers = 0

    def carve_kv(self, rows, dh):
        n = rows * dh
        k = self.arena[self.kv_off:self.kv_off + n].view(1, rows, dh // 128, 128)
        v = self.arena[self.kv_off + n:self.kv_off + 2 * n].view(1, rows, dh // 128, 128)
        return k, v

    def exchange_buffers(self, L):
        return self.q_buf[:L], self.o_heads[:L], self.o_rows[:L // self.world]

    def peer_ptrs(self, t):
        off = t.data_ptr() - self.arena.data_ptr()
        assert 0 <= off < self.arena.numel() * 2, "tensor outside the symmetric arena"
        return (ctypes.c_void_p * self.world)(*[b + off for b in self.peer_base])

    def barrier(self):
        self.barriers += 1

    def gather_rows(self, x):
        return x


@pytest.mark.parametrize("world", [2, 4, 8])
def test_p2p_exchange_pointer_arithmetic_up_to_8_ranks(on_recorder, world):
    """Every rank of a `world`-way sequence-parallel run issues launches whose shard sizes, row offsets and PER-PEER
    destination addresses are what the layout demands: rank r's rows land at row r*n of every peer's q buffer / K / V
    slot (columns = the peer's heads), and the attention output of my heads goes back to row 0.. of the owners' row
    buffers at my column block."""
    from realtime_video_b200.dit import CausalWanModel
    torch.manual_seed(1)
    heads, D = 8, 1024
    m = CausalWanModel(dim=D, ffn_dim=512, num_heads=heads, num_layers=1, text_dim=128).to(torch.bfloat16).eval()
    blk = m.blocks[0]
    blk.self_attn.fuse_projections()
    fs_, frames = 64, 3
    L = frames * fs_                   # 192 tokens
    n_loc, dh = L // world, D // world
    ca = {"k": torch.randn(1, 512, heads, 128).bfloat16(), "v": torch.randn(1, 512, heads, 128).bfloat16(), "is_init": True}
    e0 = torch.randn(frames, 6, D).bfloat16()
    for rank in range(world):
        sp = _FakeSP(rank, world, L, D)
        m.sp = sp
        k, v = sp.carve_kv(2 * L, dh)
        kv = {"k": k, "v": v, "global_end_index": L, "local_end_index": L}
        x = torch.randn(n_loc, D).bfloat16()
        on_recorder.kr_record_clear()
        with torch.no_grad():
            m._block(blk, x, e0, (frames, 8, 8), None, kv, ca, L, None)
        calls = parse(on_recorder.kr_record_dump().decode())
        m.sp = None
        fns = [c["fn"] for c in calls]
        assert fns == ["add_mod", "ln", "gemm", "qkv_post", "attn", "p2p_scatter_rows", "gemm", "ln", "gemm", "rmsnorm",
                       "attn", "gemm", "ln", "gemm", "gemm"] and sp.barriers == 2
        base = sp.arena.data_ptr()
        ln, qkv, post, attn, scat, oproj = calls[1], calls[2], calls[3], calls[4], calls[5], calls[6]
        assert ln["rows"] == n_loc and ln["row_offset"] == rank * n_loc and qkv["M"] == n_loc and qkv["N"] == 3 * D
        assert post["rows"] == n_loc and post["row_offset"] == rank * n_loc and post["peer_cols"] == dh
        assert post["ldqo"] == post["ldko"] == post["ldvo"] == dh and post["start_frame"] == L // fs_
        k_slot_off = kv["k"][0].view(-1, dh)[L:].data_ptr() - base          # slot [L, 2L) of the local cache
        v_slot_off = kv["v"][0].view(-1, dh)[L:].data_ptr() - base
        for d in range(world):
            pb = sp.peer_base[d]
            assert post[f"qp{d}"] == pb + (sp.q_buf.data_ptr() - base) + rank * n_loc * dh * 2
            assert post[f"kp{d}"] == pb + k_slot_off + rank * n_loc * dh * 2
            assert post[f"vp{d}"] == pb + v_slot_off + rank * n_loc * dh * 2
            assert scat[f"dp{d}"] == pb + (sp.o_rows.data_ptr() - base) + rank * dh * 2
        for d in range(world, 8):
            assert post[f"qp{d}"] == 0
        assert attn["Lq"] == L and attn["Lkv"] == 2 * L and attn["heads"] == heads // world
        assert attn["q"] == sp.q_buf.data_ptr() and attn["ldq"] == dh and attn["out"] == sp.o_heads.data_ptr()
        assert scat["src"] == sp.o_heads.data_ptr() and scat["rows"] == L and scat["rows_per_peer"] == n_loc
        assert scat["cols"] == dh and scat["ld_dst"] == D and scat["world"] == world
        assert oproj["a"] == sp.o_rows.data_ptr() and oproj["M"] == n_loc and oproj["row_offset"] == rank * n_loc
        assert (kv["global_end_index"], kv["local_end_index"]) == (2 * L, 2 * L)
