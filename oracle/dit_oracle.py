"""TEST INFRASTRUCTURE — CPU oracle for the causal Wan DiT hot path (not shipped, not measured).

A plain-torch restatement of the reference's algorithm, written from its behaviour; every
function cites the reference file:line it follows (paths relative to krea-ai/realtime-video).
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` leg may import this module.  The product path (``realtime_video_b200``) never does.

Pinning: ``tests/golden/make_dit_goldens.py`` imports the UNMODIFIED reference modules in the
build container (CPU, import shims only) and stores inputs/outputs under ``tests/golden/``;
``tests/test_oracle_dit.py`` checks this oracle against those fixtures (fp32: ~1e-5 relative).

The oracle computes in the dtype of the parameters it is given: fp32 parameters give the exact
restatement; bf16 parameters reproduce the reference's eager-bf16 rounding points because the
op order below is the reference's op order.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F


# ---------------------------------------------------------------------------------------------
# embeddings / rotary tables
# ---------------------------------------------------------------------------------------------
def sinusoidal_embedding_1d(dim: int, position: torch.Tensor) -> torch.Tensor:
    """wan/modules/model.py:15-24 — cat(cos, sin) of t * 10000^(-i/half), float64."""
    half = dim // 2
    position = position.to(torch.float64)
    freqs = torch.pow(10000, -torch.arange(half, dtype=torch.float64, device=position.device).div(half))
    s = torch.outer(position, freqs)
    return torch.cat([torch.cos(s), torch.sin(s)], dim=1)


def rope_angles(max_seq_len: int, dim: int, theta: float = 10000.0) -> torch.Tensor:
    """Angles of wan/modules/model.py:28-35 (the reference stores polar(1, angle))."""
    return torch.outer(torch.arange(max_seq_len, dtype=torch.float64),
                       1.0 / torch.pow(theta, torch.arange(0, dim, 2, dtype=torch.float64).div(dim)))


def rope_table(head_dim: int, max_len: int = 1024) -> torch.Tensor:
    """causal_model.py:638-645 — angle table [max_len, head_dim/2] = cat(t, h, w) (float64)."""
    d = head_dim
    return torch.cat([rope_angles(max_len, d - 4 * (d // 6)),
                      rope_angles(max_len, 2 * (d // 6)),
                      rope_angles(max_len, 2 * (d // 6))], dim=1)


def rope_apply(x: torch.Tensor, grid, angles: torch.Tensor, start_frame: int = 0) -> torch.Tensor:
    """causal_model.py:143-171 (start_frame) / model.py:39-66 (start_frame = 0).

    x: [L, n, d]; grid = (f, h, w); angles [1024, d/2] float64.  Adjacent pairs (2i, 2i+1) are
    complex numbers rotated by the angle of frame / row / column position; float64 math,
    result cast back to x.dtype."""
    f, h, w = grid
    L, n, d = x.shape
    c = d // 2
    seq = f * h * w
    ct, ch = c - 2 * (c // 3), c // 3
    a = angles.to(x.device)
    ang = torch.cat([
        a[start_frame:start_frame + f, :ct].view(f, 1, 1, -1).expand(f, h, w, -1),
        a[:h, ct:ct + ch].view(1, h, 1, -1).expand(f, h, w, -1),
        a[:w, ct + ch:].view(1, 1, w, -1).expand(f, h, w, -1)], dim=-1).reshape(seq, 1, c)
    xr = x[:seq].to(torch.float64).reshape(seq, n, c, 2)
    cs, sn = torch.cos(ang), torch.sin(ang)
    x0, x1 = xr[..., 0], xr[..., 1]
    out = torch.stack([x0 * cs - x1 * sn, x0 * sn + x1 * cs], dim=-1).reshape(seq, n, d)
    out = torch.cat([out, x[seq:].to(torch.float64)])
    return out.to(x.dtype)


# ---------------------------------------------------------------------------------------------
# norms / attention
# ---------------------------------------------------------------------------------------------
def rms_norm(x: torch.Tensor, weight: torch.Tensor, eps: float) -> torch.Tensor:
    """wan/modules/model.py:69-85 — fp32 normalise, cast back, then multiply by weight."""
    xf = x.float()
    return (xf * torch.rsqrt(xf.pow(2).mean(dim=-1, keepdim=True) + eps)).type_as(x) * weight


def layer_norm(x: torch.Tensor, eps: float, weight=None, bias=None) -> torch.Tensor:
    """wan/modules/model.py:88-98."""
    return F.layer_norm(x, (x.shape[-1],), weight, bias, eps).type_as(x)


def block_causal_mask(Lq: int, Lkv: int, block_len: int, window: int = 0, device=None) -> torch.Tensor:
    """causal_model.py:109-141 get_block_mask rule: allowed iff kv < ends[q] (| q == kv), with
    optional local window kv >= ends[q] - window."""
    qi = torch.arange(Lq, device=device)[:, None]
    ki = torch.arange(Lkv, device=device)[None, :]
    ends = (qi // block_len + 1) * block_len
    m = ki < ends
    if window > 0:
        m = m & (ki >= ends - window)
    return m | (qi == ki)


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, mask: Optional[torch.Tensor] = None):
    """softmax(q k^T / sqrt(d)) v over [L, n, d] tensors — what flash_attn_func(q, k, v)
    computes (wan/modules/attention.py:65-70): no mask, default scale; fp32 softmax."""
    d = q.shape[-1]
    qf, kf, vf = (t.float().transpose(0, 1) for t in (q, k, v))     # [n, L, d]
    s = qf @ kf.transpose(1, 2) / math.sqrt(d)
    if mask is not None:
        s = s.masked_fill(~mask, float("-inf"))
    p = torch.softmax(s, dim=-1)
    return (p @ vf).transpose(0, 1).to(q.dtype)


# ---------------------------------------------------------------------------------------------
# the model
# ---------------------------------------------------------------------------------------------
class DiTConfig:
    def __init__(self, dim=5120, ffn_dim=13824, num_heads=40, num_layers=40, in_dim=16, out_dim=16,
                 freq_dim=256, text_dim=4096, text_len=512, eps=1e-6, local_attn_size=-1,
                 sink_size=0, patch_size=(1, 2, 2), frame_seqlen_const=1560, cross_attn_norm=True):
        self.dim, self.ffn_dim, self.num_heads, self.num_layers = dim, ffn_dim, num_heads, num_layers
        self.in_dim, self.out_dim, self.freq_dim = in_dim, out_dim, freq_dim
        self.text_dim, self.text_len, self.eps = text_dim, text_len, eps
        self.local_attn_size, self.sink_size, self.patch_size = local_attn_size, sink_size, patch_size
        # the reference hard-codes 1560 tokens per latent frame (causal_model.py:192, :351)
        self.frame_seqlen_const = frame_seqlen_const
        self.head_dim = dim // num_heads
        # norm3 = WanLayerNorm(affine) if cross_attn_norm else nn.Identity() (causal_model.py:424-426); True for every Wan 2.1 model
        self.cross_attn_norm = cross_attn_norm


def new_kv_cache(cfg: DiTConfig, size: int, dtype, device="cpu") -> List[Dict]:
    """pipeline/causal_inference.py:279-310."""
    return [{"k": torch.zeros(1, size, cfg.num_heads, cfg.head_dim, dtype=dtype, device=device),
             "v": torch.zeros(1, size, cfg.num_heads, cfg.head_dim, dtype=dtype, device=device),
             "global_end_index": 0, "local_end_index": 0} for _ in range(cfg.num_layers)]


def new_crossattn_cache(cfg: DiTConfig, dtype, device="cpu") -> List[Dict]:
    """pipeline/causal_inference.py:312-339."""
    return [{"k": torch.zeros(1, cfg.text_len, cfg.num_heads, cfg.head_dim, dtype=dtype, device=device),
             "v": torch.zeros(1, cfg.text_len, cfg.num_heads, cfg.head_dim, dtype=dtype, device=device),
             "is_init": False} for _ in range(cfg.num_layers)]


class DiTOracle:
    """Functional CausalWanModel._forward_inference over a reference-keyed state dict
    (keys as in CausalWanModel.state_dict(): 'blocks.0.self_attn.q.weight', ...)."""

    def __init__(self, cfg: DiTConfig, params: Dict[str, torch.Tensor]):
        self.cfg = cfg
        self.p = params
        self.angles = rope_table(cfg.head_dim)
        self.max_attention_size = 32760 if cfg.local_attn_size == -1 else \
            cfg.local_attn_size * cfg.frame_seqlen_const      # causal_model.py:192

    def lin(self, name: str, x: torch.Tensor) -> torch.Tensor:
        return F.linear(x, self.p[name + ".weight"], self.p.get(name + ".bias"))

    # -- causal_model.py:218-397 ---------------------------------------------------------
    def self_attn(self, i: int, x, grid, kv_cache, current_start, mask_args):
        cfg, pre = self.cfg, f"blocks.{i}.self_attn."
        L = x.shape[0]
        n, d = cfg.num_heads, cfg.head_dim
        q = rms_norm(self.lin(pre + "q", x), self.p[pre + "norm_q.weight"], cfg.eps).view(L, n, d)
        k = rms_norm(self.lin(pre + "k", x), self.p[pre + "norm_k.weight"], cfg.eps).view(L, n, d)
        v = self.lin(pre + "v", x).view(L, n, d)
        if mask_args is not None:
            # recompute / FlexAttention branch, causal_model.py:305-348
            rq = rope_apply(q, grid, self.angles).type_as(v)
            rk = rope_apply(k, grid, self.angles).type_as(v)
            kv_cache["k"][0, :L] = rk
            kv_cache["v"][0, :L] = v
            kv_cache["global_end_index"] = L
            kv_cache["local_end_index"] = L
            # :316-348 — q/k/v are right-padded with ZERO rows to a multiple of 128 and the mask is
            # built over the padded length, so queries of an incomplete last block also see the
            # padded keys (score 0, value 0); the padded query rows are sliced away.
            pad = math.ceil(L / 128) * 128 - L
            zp = rk.new_zeros(pad, n, d)
            mask = block_causal_mask(L, L + pad, mask_args["block_len"], mask_args.get("window", 0), x.device)
            out = attention(rq, torch.cat([rk, zp]), torch.cat([v, zp]), mask)
        else:
            fs = cfg.frame_seqlen_const
            start_frame = current_start // fs                              # :351-352
            rq = rope_apply(q, grid, self.angles, start_frame).type_as(v)
            rk = rope_apply(k, grid, self.angles, start_frame).type_as(v)
            current_end = current_start + L
            sink_tokens = cfg.sink_size * fs
            kv_size = kv_cache["k"].shape[1]
            g_end, l_end = int(kv_cache["global_end_index"]), int(kv_cache["local_end_index"])
            if cfg.local_attn_size != -1 and current_end > g_end and L + l_end > kv_size:
                evicted = L + l_end - kv_size                              # :363-379
                rolled = l_end - evicted - sink_tokens
                for key in ("k", "v"):
                    src = kv_cache[key][:, sink_tokens + evicted:sink_tokens + evicted + rolled].clone()
                    kv_cache[key][:, sink_tokens:sink_tokens + rolled] = src
                local_end = l_end + current_end - g_end - evicted
            else:
                local_end = l_end + current_end - g_end                    # :380-385
            local_start = local_end - L
            kv_cache["k"][0, local_start:local_end] = rk
            kv_cache["v"][0, local_start:local_end] = v
            lo = max(0, local_end - self.max_attention_size)               # :386-390
            out = attention(rq, kv_cache["k"][0, lo:local_end], kv_cache["v"][0, lo:local_end])
            kv_cache["global_end_index"] = current_end
            kv_cache["local_end_index"] = local_end
        return self.lin(pre + "o", out.reshape(L, n * d))

    # -- wan/modules/model.py:171-228 -----------------------------------------------------
    def cross_attn(self, i: int, x, context, cache):
        cfg, pre = self.cfg, f"blocks.{i}.cross_attn."
        n, d = cfg.num_heads, cfg.head_dim
        L = x.shape[0]
        q = rms_norm(self.lin(pre + "q", x), self.p[pre + "norm_q.weight"], cfg.eps).view(L, n, d)
        if cache is not None and cache["is_init"]:
            k, v = cache["k"][0], cache["v"][0]
        else:
            k = rms_norm(self.lin(pre + "k", context), self.p[pre + "norm_k.weight"], cfg.eps).view(-1, n, d)
            v = self.lin(pre + "v", context).view(-1, n, d)
            if cache is not None:
                cache["is_init"] = True
                cache["k"], cache["v"] = k[None], v[None]
        out = attention(q, k, v)
        return self.lin(pre + "o", out.reshape(L, n * d))

    # -- causal_model.py:440-492 ----------------------------------------------------------
    def block(self, i: int, x, e0, grid, context, kv_cache, crossattn_cache, current_start, mask_args):
        cfg, pre = self.cfg, f"blocks.{i}."
        Fr = e0.shape[0]
        fs = x.shape[0] // Fr
        e = (self.p[pre + "modulation"] + e0).chunk(6, dim=1)              # each [F, 1, D]

        def per_frame(t):
            return t.unflatten(0, (Fr, fs))

        h = (per_frame(layer_norm(x, cfg.eps)) * (1 + e[1]) + e[0]).flatten(0, 1)
        y = self.self_attn(i, h, grid, kv_cache, current_start, mask_args)
        x = x + (per_frame(y) * e[2]).flatten(0, 1)
        h = layer_norm(x, cfg.eps, self.p[pre + "norm3.weight"], self.p[pre + "norm3.bias"]) if cfg.cross_attn_norm else x
        x = x + self.cross_attn(i, h, context, crossattn_cache)
        h = (per_frame(layer_norm(x, cfg.eps)) * (1 + e[4]) + e[3]).flatten(0, 1)
        y = self.lin(pre + "ffn.2", F.gelu(self.lin(pre + "ffn.0", h), approximate="tanh"))
        x = x + (per_frame(y) * e[5]).flatten(0, 1)
        return x

    # -- causal_model.py:825-954 ----------------------------------------------------------
    def forward_inference(self, x, t, context, kv_cache, crossattn_cache, current_start=0,
                          mask_args=None, return_hidden=False):
        """x [C, F, H, W]; t [F] timesteps; context [<=text_len, text_dim] -> flow [C, F, H, W]."""
        cfg = self.cfg
        C, Fr, H, W = x.shape
        dt = self.p["patch_embedding.weight"].dtype
        tok = F.conv3d(x[None].to(dt), self.p["patch_embedding.weight"], self.p["patch_embedding.bias"],
                       stride=cfg.patch_size)                              # :874
        grid = tuple(tok.shape[2:])
        tok = tok.flatten(2).transpose(1, 2)[0]                            # [L, D]
        emb = sinusoidal_embedding_1d(cfg.freq_dim, t.flatten()).to(dt)    # :890
        e = self.lin("time_embedding.2", F.silu(self.lin("time_embedding.0", emb)))       # [F, D]
        e0 = self.lin("time_projection.1", F.silu(e)).unflatten(1, (6, cfg.dim))          # [F, 6, D]
        ctx = torch.cat([context, context.new_zeros(cfg.text_len - context.shape[0], context.shape[1])])
        ctx = self.lin("text_embedding.2", F.gelu(self.lin("text_embedding.0", ctx.to(dt)), approximate="tanh"))
        hidden = []
        for i in range(cfg.num_layers):
            tok = self.block(i, tok, e0, grid, ctx, kv_cache[i],
                             crossattn_cache[i] if crossattn_cache is not None else None,
                             current_start, mask_args)
            if return_hidden:
                hidden.append(tok)
        # head, causal_model.py:512-523 (e is the pre-projection time embedding, :951)
        fs = tok.shape[0] // Fr
        eh = (self.p["head.modulation"] + e[:, None, :]).chunk(2, dim=1)   # [F,1,D] x2
        h = layer_norm(tok, cfg.eps).unflatten(0, (Fr, fs)) * (1 + eh[1]) + eh[0]
        out = self.lin("head.head", h)                                     # [F, fs, 4*C_out]
        # unpatchify, causal_model.py:1126-1149
        f, hh, ww = grid
        u = out.reshape(f, hh, ww, *cfg.patch_size, cfg.out_dim)
        u = torch.einsum("fhwpqrc->cfphqwr", u).reshape(cfg.out_dim, f * cfg.patch_size[0],
                                                        hh * cfg.patch_size[1], ww * cfg.patch_size[2])
        return (u, hidden) if return_hidden else u


# ---------------------------------------------------------------------------------------------
# scheduler / wrapper arithmetic
# ---------------------------------------------------------------------------------------------
class FlowMatchSchedulerOracle:
    """utils/scheduler.py:106-176 with shift, sigma_min=0, extra_one_step=True, 1000 steps."""

    def __init__(self, shift: float = 5.0, num_train_timesteps: int = 1000):
        sig = torch.linspace(1.0, 0.0, num_train_timesteps + 1)[:-1]
        self.sigmas = shift * sig / (1 + (shift - 1) * sig)
        self.timesteps = self.sigmas * num_train_timesteps

    def sigma_of(self, timestep: torch.Tensor) -> torch.Tensor:
        idx = torch.argmin((self.timesteps.to(timestep.device).unsqueeze(0) -
                            timestep.unsqueeze(1)).abs(), dim=1)
        return self.sigmas.to(timestep.device)[idx]

    def add_noise(self, x0, noise, timestep):
        """scheduler.py:159-176."""
        sigma = self.sigma_of(timestep.flatten()).reshape(-1, 1, 1, 1)
        return ((1 - sigma) * x0 + sigma * noise).type_as(noise)


def flow_to_x0(flow, xt, timestep, sched: FlowMatchSchedulerOracle):
    """utils/wan_wrapper.py:181-205 — x0 = xt - sigma_t * flow in float64. [F, C, H, W] inputs."""
    ts = sched.timesteps.double().to(flow.device)
    idx = torch.argmin((ts.unsqueeze(0) - timestep.double().unsqueeze(1)).abs(), dim=1)
    sigma = sched.sigmas.double().to(flow.device)[idx].reshape(-1, 1, 1, 1)
    return (xt.double() - sigma * flow.double()).to(flow.dtype)


def denoising_schedule(sched: FlowMatchSchedulerOracle, strength: float = 1.0, steps: int = 4):
    """v2v.py:133-136 over release_server.py:561 zero-padded timesteps."""
    zp = torch.cat((sched.timesteps, torch.tensor([0], dtype=torch.float32)))
    lst = torch.linspace(strength * 1000, 0, steps, dtype=torch.float32).to(torch.long)
    return zp[1000 - lst]
