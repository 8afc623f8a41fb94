"""Parity at the REAL geometry of the BASELINE configs (the other GPU tests use a dim-256 / 96-token fixture model):

  C2  one Wan-14B-dim DiT layer (d 5120, 40 heads, ffn 13824), Lq 4680 against Lkv 9360: CUDA path vs the fp32 oracle
      on the host; the attention kernel alone at its 760-CTA bench grid (and the recompute shape) vs fp32 torch;
  C1  one Wan-1.3B-dim layer (d 1536, 12 heads, ffn 8960) at 320x192 (240 tokens/frame);
  C4  3600 tokens/frame (1280x720), kv_cache_num_frames=5: recompute branch over 5 context frames, then the cache
      branch with Lq 10800 / Lkv 28800 (small width, the oracle with frame_seqlen_const=3600);
  VAE the server decoder at latent 60x104 -> 480x832 px against a reference-executed golden (tests/golden/vae_fullres.npz).

Tolerances: bf16 kernels vs fp32 references, one layer: rel-L2 <= 1.5e-2 (the reference's own bf16-vs-fp32 gap is 5.9e-3,
SURVEY.md §8a); attention rel-L2 <= 1e-2; VAE (fp16 storage, ~40 convs deep) mean|d| <= 6e-3, rel-L2 <= 2e-2."""
import math

import pytest
import torch

from tests.golden_io import GOLDEN, load_npz, rel_l2

pytestmark = pytest.mark.gpu


def _attn_ref_fp32(q, k, v, heads, block_len=0, pad_keys=0):
    """Per-head fp32 softmax(QK^T/sqrt(d))V; block-causal rule + the padded-key quirk of the flex branch
    (zero keys appended up to a multiple of 128 are visible to the LAST block's queries, causal_model.py:316-348)."""
    Lq, Lkv = q.shape[0], k.shape[0]
    out = torch.empty(Lq, heads * 128, dtype=torch.float32, device=q.device)
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        for h in range(heads):
            cs = slice(h * 128, (h + 1) * 128)
            s = q[:, cs].float() @ k[:, cs].float().t() / math.sqrt(128)
            vv = v[:, cs].float()
            if block_len:
                qi = torch.arange(Lq, device=q.device)[:, None]
                ki = torch.arange(Lkv + pad_keys, device=q.device)[None, :]
                if pad_keys:
                    s = torch.cat([s, s.new_zeros(Lq, pad_keys)], dim=1)
                    vv = torch.cat([vv, vv.new_zeros(pad_keys, 128)])
                s = s.masked_fill(~(ki < (qi // block_len + 1) * block_len), float("-inf"))
            out[:, cs] = torch.softmax(s, dim=-1) @ vv
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    return out


@pytest.mark.parametrize("Lq,Lkv,heads,block_len", [
    (4680, 9360, 40, 0),          # C2 denoise pass: 19 q-tiles x 40 heads = 760 CTAs, short last CTA per head
    (4680, 4680, 40, 4680),       # C2 recompute pass: block-causal rule (one 3-frame block) + 56 padded keys
    (10800, 28800, 8, 0),         # C4 denoise pass geometry (8 of the 40 heads)
])
def test_attention_kernel_at_the_bench_shapes(Lq, Lkv, heads, block_len):
    from realtime_video_b200 import ops
    torch.manual_seed(Lq + Lkv)
    D = heads * 128
    q = torch.randn(Lq, D, device="cuda").bfloat16()
    k = torch.randn(Lkv, D, device="cuda").bfloat16()
    v = torch.randn(Lkv, D, device="cuda").bfloat16()
    pad = (math.ceil(Lkv / 128) * 128 - Lkv) if block_len else 0
    out = ops.attention(q, k, v, heads=heads, block_len=block_len, pad_keys=pad)
    r = rel_l2(out.float(), _attn_ref_fp32(q, k, v, heads, block_len, pad))
    assert r < 1e-2, r


def _one_layer_vs_oracle(dims, latent_hw, kv_frames=3, fs_const=None, recompute_frames=0, tol=1.5e-2):
    """Product CausalWanModel (1 layer) on the GPU vs oracle.DiTOracle in fp32 on the host, same seeded weights:
    optional recompute pass over ``recompute_frames`` context frames (block-causal branch), otherwise a first
    cache-branch block; then a cache-branch block of 3 frames behind it."""
    from oracle import dit_oracle as O
    from realtime_video_b200.dit import CausalWanModel
    H, W = latent_hw
    fs = (H // 2) * (W // 2)
    torch.manual_seed(0)
    with torch.device("cuda"):
        m = CausalWanModel(num_layers=1, **dims)
    with torch.no_grad():
        m.head.head.weight.normal_(std=0.02)
        for n_, p_ in m.named_parameters():
            if n_.endswith(".bias"):
                p_.normal_(std=0.02)
    m = m.to(torch.bfloat16).eval()
    m.blocks[0].self_attn.fuse_projections()
    m.num_frame_per_block = 3
    sd = {k: v.float().cpu() for k, v in m.state_dict().items() if "to_qkv" not in k}
    cfg = O.DiTConfig(num_layers=1, frame_seqlen_const=fs_const or fs, **dims)
    orc = O.DiTOracle(cfg, sd)
    n_ctx = recompute_frames or 3
    rows = (n_ctx + 3) * fs
    n, d = dims["num_heads"], 128
    kv = [dict(k=torch.zeros(1, rows, n, d, dtype=torch.bfloat16, device="cuda"),
               v=torch.zeros(1, rows, n, d, dtype=torch.bfloat16, device="cuda"), global_end_index=0, local_end_index=0)]
    ca = [dict(k=torch.zeros(1, 512, n, d, dtype=torch.bfloat16, device="cuda"),
               v=torch.zeros(1, 512, n, d, dtype=torch.bfloat16, device="cuda"), is_init=False)]
    kv_o, ca_o = O.new_kv_cache(cfg, rows, torch.float32), O.new_crossattn_cache(cfg, torch.float32)
    g = torch.Generator().manual_seed(3)
    ctx = torch.randn(40, dims.get("text_dim", 4096), generator=g).bfloat16()
    x_ctx = torch.randn(16, n_ctx, H, W, generator=g).bfloat16()
    x_new = torch.randn(16, 3, H, W, generator=g).bfloat16()
    res = {}
    with torch.no_grad():
        if recompute_frames:
            m.block_mask = m._prepare_blockwise_causal_attn_mask("cuda", num_frames=n_ctx, frame_seqlen=fs,
                                                                 num_frame_per_block=3, local_attn_size=-1)
            mask_args = dict(block_len=3 * fs)
        else:
            mask_args = None
        t0 = torch.zeros(n_ctx) if recompute_frames else torch.full((n_ctx,), 1000.0)
        got = m(x_ctx.cuda()[None], t=t0.cuda()[None], context=ctx.cuda()[None], seq_len=1 << 30, kv_cache=kv,
                crossattn_cache=ca, current_start=n_ctx * fs if recompute_frames else 0)[0]
        m.block_mask = None
        ref = orc.forward_inference(x_ctx.float(), t0, ctx.float(), kv_o, ca_o,
                                    n_ctx * fs if recompute_frames else 0, mask_args=mask_args)
        res["first"] = rel_l2(got.float().cpu(), ref)
        t1 = torch.full((3,), 750.0)
        got = m(x_new.cuda()[None], t=t1.cuda()[None], context=ctx.cuda()[None], seq_len=1 << 30, kv_cache=kv,
                crossattn_cache=ca, current_start=n_ctx * fs)[0]
        ref = orc.forward_inference(x_new.float(), t1, ctx.float(), kv_o, ca_o, n_ctx * fs)
        res["second"] = rel_l2(got.float().cpu(), ref)
    assert kv[0]["local_end_index"] == rows and kv_o[0]["local_end_index"] == rows
    rk = rel_l2(kv[0]["k"][0].float().cpu(), kv_o[0]["k"][0])
    assert res["first"] < tol and res["second"] < tol and rk < tol, (res, rk)


def test_c2_one_14b_dim_layer_vs_oracle():
    """d 5120 / 40 heads / ffn 13824 at 60x104 latents: block 0 (Lq = Lkv = 4680), then Lq 4680 vs Lkv 9360."""
    _one_layer_vs_oracle(dict(dim=5120, ffn_dim=13824, num_heads=40, text_dim=4096), (60, 104))


def test_c1_one_1p3b_dim_layer_vs_oracle():
    """Wan2.1-T2V-1.3B dims (wan/configs/wan_t2v_1_3B.py:21-28) at 320x192 -> latent 24x40, 240 tokens/frame."""
    _one_layer_vs_oracle(dict(dim=1536, ffn_dim=8960, num_heads=12, text_dim=4096), (24, 40))


def test_c4_720p_recompute_and_cache_branch_vs_oracle():
    """1280x720 -> latent 90x160, 3600 tokens/frame, kv_cache_num_frames=5: block-causal recompute over 5 context
    frames (18000 tokens, padded-key quirk) then 3 new frames against 28800 cached keys."""
    _one_layer_vs_oracle(dict(dim=256, ffn_dim=512, num_heads=2, text_dim=128), (90, 160), recompute_frames=5,
                         fs_const=3600)


@pytest.mark.skipif(not (GOLDEN / "vae_fullres.npz").exists(), reason="vae_fullres.npz not generated")
def test_vae_decoder_at_480x832_vs_reference_golden():
    """Exercises the 16x8-pixel halo tiles, the 60x104 -> 480x832 upsampling chain and the steady-block chunking at
    the bench geometry; golden = unmodified reference VAEDecoderWrapper, fp32, every 4th pixel."""
    from realtime_video_b200.factory import synthetic_vae_decoder
    g = load_npz("vae_fullres.npz")
    dec = synthetic_vae_decoder(device="cuda")
    cache = [None] * 55
    with torch.no_grad():
        for call, frames in ((0, 9), (1, 12)):
            px, cache = dec(g[f"z{call}"].cuda().half(), *cache)
            assert px.shape == (1, frames, 3, 480, 832)
            got = px[0, :, :, 1::4, 1::4].float().cpu()
            ref = g[f"px{call}_sub"].float()
            mad = (got - ref).abs().mean().item()
            r = rel_l2(got, ref)
            assert mad < 6e-3 and r < 2e-2, (call, mad, r)


def test_wan_vae_wrapper_cached_decode_streams_like_one_shot():
    """WanVAEWrapper.decode_to_pixel(use_cache=True) = WanVAE_.cached_decode (wan/modules/vae.py:545-567): the feature
    cache survives between calls, so two cached calls equal one uncached call over the concatenated latents, and match
    the reference streaming golden."""
    from realtime_video_b200.factory import synthetic_vae_params
    from realtime_video_b200.wan_wrapper import WanVAEWrapper
    g = load_npz("vae_small.npz")
    vae = WanVAEWrapper(load_pretrained=False)
    vae.model.load_state_dict(synthetic_vae_params(seed=0), strict=False)
    vae = vae.to(device="cuda", dtype=torch.bfloat16).eval()
    z0, z1 = g["s8x12/z0"].cuda().bfloat16(), g["s8x12/z1"].cuda().bfloat16()
    with torch.no_grad():
        a0 = vae.decode_to_pixel(z0, use_cache=True)
        a1 = vae.decode_to_pixel(z1, use_cache=True)
        both = vae.decode_to_pixel(torch.cat([z0, z1], dim=1), use_cache=False)
    assert a0.shape == (1, 9, 3, 64, 96) and a1.shape == (1, 12, 3, 64, 96)
    assert torch.equal(torch.cat([a0, a1], dim=1), both)
    for got, name in ((a0, "s8x12/px0"), (a1, "s8x12/px1")):
        assert (got.float().cpu() - g[name]).abs().mean().item() < 1.2e-2      # bf16 storage on the classic path
