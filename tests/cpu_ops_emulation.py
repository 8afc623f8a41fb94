"""fp32 torch stand-ins for the ``realtime_video_b200.ops`` functions the DiT host schedule calls — TEST
INFRASTRUCTURE: lets ``-m "not gpu"`` tests execute realtime_video_b200/dit.py's HOST logic (cache-slot
arithmetic, rolling eviction, split QKV output, block-mask / padded-key bookkeeping, cross-attention cache,
per-frame modulation indexing, unpatchify) on the CPU and compare it with the reference goldens.  Each function
implements the documented contract of the C-ABI entry point of the same name (include/krea_b200.h) without the
16-bit rounding points; nothing here is reachable from the product path (tests monkeypatch ``dit.ops``)."""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn.functional as F

from oracle import dit_oracle as O

EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_GATE_RES, EPI_BIAS_RES, EPI_F32, EPI_MUL = 0, 1, 2, 3, 4, 5
launch_count = 0


def _store(out: Optional[torch.Tensor], val: torch.Tensor) -> torch.Tensor:
    if out is None:
        return val
    out.copy_(val)
    return out


def gemm(a, w, bias=None, *, epilogue=EPI_BIAS, out=None, residual=None, gate=None, rows_per_gate=0, alpha=1.0,
         out2=None, n_split=0, row_offset=0):
    y = a.float() @ w.float().t()
    if bias is not None:
        y = y + bias.float()
    if epilogue == EPI_BIAS_GELU:
        y = F.gelu(y, approximate="tanh")
    elif epilogue == EPI_BIAS_GATE_RES:
        rows = torch.arange(a.shape[0]) + row_offset
        y = residual.float() + y * gate.float()[rows // rows_per_gate]
    elif epilogue == EPI_BIAS_RES:
        y = residual.float() + y
    elif epilogue == EPI_MUL:
        y = residual.float() * y
    elif epilogue == EPI_F32:
        return _store(out, y * alpha)
    y = y.to(a.dtype)
    if out2 is not None:
        out2.copy_(y[:, n_split:])
        y = y[:, :n_split]
    return _store(out, y)


def attention(q, k, v, *, heads, out=None, softmax_scale=None, block_len=0, window=0, pad_keys=0):
    Lq, Lkv = q.shape[0], k.shape[0]
    d = q.shape[-1] // heads
    scale = softmax_scale if softmax_scale is not None else 1.0 / math.sqrt(d)
    qf = q.float().reshape(Lq, heads, d).transpose(0, 1)
    kf = k.float().reshape(Lkv, heads, d).transpose(0, 1)
    vf = v.float().reshape(Lkv, heads, d).transpose(0, 1)
    s = qf @ kf.transpose(1, 2) * scale                                   # [heads, Lq, Lkv]
    n_ph = torch.zeros(Lq)
    if block_len > 0:
        qi = torch.arange(Lq)[:, None]
        ki = torch.arange(Lkv)[None, :]
        hi = (qi // block_len + 1) * block_len
        ok = ki < hi
        if window > 0:
            ok = ok & (ki >= hi - window)
        s = s.masked_fill(~ok, float("-inf"))
        # zero-padded phantom keys of the FlexAttention path (score 0, value 0) for rows whose block runs past Lkv
        n_ph = (hi[:, 0] - Lkv).clamp(min=0, max=pad_keys).float()
    m = torch.maximum(s.amax(dim=-1), torch.where(n_ph > 0, 0.0, float("-inf"))[None])
    p = torch.exp(s - m[..., None])
    den = p.sum(-1) + n_ph[None] * torch.exp(-m)
    o = (p @ vf) / den[..., None]
    return _store(out, o.transpose(0, 1).reshape(Lq, heads * d).to(q.dtype))


def ln_modulate(x, *, eps, weight=None, bias=None, mod=None, shift_idx=0, scale_idx=1, rows_per_frame=0, out=None,
                row_offset=0):
    y = F.layer_norm(x.float(), (x.shape[-1],), None if weight is None else weight.float(),
                     None if bias is None else bias.float(), eps)
    if mod is not None:
        fr = (torch.arange(x.shape[0]) + row_offset) // rows_per_frame
        y = y * (1 + mod.float()[fr, scale_idx]) + mod.float()[fr, shift_idx]
    return _store(out, y.to(x.dtype))


def rmsnorm(x, weight, eps, out=None):
    xf = x.float()
    return _store(out, (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps) * weight.float()).to(x.dtype))


def qkv_norm_rope(q, k, v, wq, wk, q_out, k_out, v_out, rope, *, head_dim, grid_h, grid_w, start_frame, eps,
                  row_offset=0):
    """RMSNorm over all D channels, then RoPE per head: pair j of a head rotates by the angle of the token's
    frame (first c - 2(c//3) pairs), row (next c//3) or column (last c//3) position; the token of local row r is
    global row r + row_offset of the [frames, grid_h, grid_w] raster (causal_model.py:143-171)."""
    rows, D = q.shape
    n, c = D // head_dim, head_dim // 2
    fs = grid_h * grid_w
    r = torch.arange(rows) + row_offset
    tab = O.rope_table(head_dim)                                   # float64 [1024, c]
    ct, ch = c - 2 * (c // 3), c // 3
    ang = torch.cat([tab[r // fs + start_frame, :ct], tab[(r % fs) // grid_w, ct:ct + ch],
                     tab[r % grid_w, ct + ch:]], dim=-1)[:, None, :]          # [rows, 1, c]
    cs, sn = torch.cos(ang), torch.sin(ang)
    for src, w, dst in ((q, wq, q_out), (k, wk, k_out)):
        y = rmsnorm(src, w, eps).double().reshape(rows, n, c, 2)
        y0, y1 = y[..., 0], y[..., 1]
        dst.copy_(torch.stack([y0 * cs - y1 * sn, y0 * sn + y1 * cs], dim=-1).reshape(rows, D).to(q.dtype))
    if v is not None:
        v_out.copy_(v)


def add_modulation(modulation, e0, out=None):
    return _store(out, (modulation.float() + e0.float()).to(e0.dtype))


def activation(x, kind):
    return (F.silu(x.float()) if kind == "silu" else F.gelu(x.float(), approximate="tanh")).to(x.dtype)


def patchify(x):
    """[C, F, H, W] -> [F*(H/2)*(W/2), C*4]: Conv3d(kernel = stride = (1, 2, 2)) im2col, weight.view(dim, -1)
    column order (c, kh, kw)."""
    C, Fr, H, W = x.shape
    p = x.reshape(C, Fr, H // 2, 2, W // 2, 2).permute(1, 2, 4, 0, 3, 5)
    return p.reshape(Fr * (H // 2) * (W // 2), C * 4).contiguous()


def unpatchify_x0(head_out, xt, sigma, C, Fr, H, W):
    """causal_model.py:1126-1149: head columns are (kh, kw, c); x0 = xt - sigma_f * flow in float64."""
    u = head_out.reshape(Fr, H // 2, W // 2, 2, 2, C)
    flow = u.permute(0, 5, 1, 3, 2, 4).reshape(Fr, C, H, W).contiguous()
    x0 = None
    if xt is not None:
        x0 = (xt.double() - sigma.double().view(-1, 1, 1, 1) * flow.double()).to(flow.dtype)
    return flow, x0


# ---------------------------------------------------------------------------------------------
# VAE ops (channels-last activations [frames, H, W, C]); contracts: include/krea_b200.h kr_vae_*
# ---------------------------------------------------------------------------------------------
def vae_scale_input(z, mean, inv_std, w2, b2, out):
    """z [T, 16, H, W] -> x = z / inv_std + mean; y = conv2_1x1x1(x); channels-last, zero-padded to out's C."""
    x = z.float() / inv_std.float().view(1, -1, 1, 1) + mean.float().view(1, -1, 1, 1)
    y = torch.einsum("oc,tchw->thwo", w2.float(), x) + b2.float()
    out.zero_()
    out[..., :y.shape[-1]] = y.to(out.dtype)
    return out


def _rms(y, gamma, silu=True):
    C = y.shape[-1]
    o = F.normalize(y, dim=-1) * math.sqrt(C) * gamma.float()[:C]
    return F.silu(o) if silu else o


def vae_conv(x, weight, bias, *, n, cout, T, taps, tile, out_raw=None, out_norm=None, gamma=None, residual=None,
             out_pix=None, raw_frame_stride=None, sub2=False):
    """Causal conv as the kernel defines it: x holds kt-1 history frames in front of the T new ones, weight is
    [rows, kt*kh*kw*cin] with (kt, kh, kw, cin) column order, spatial zero padding, then the fused epilogue."""
    t_in, H, W, cin = x.shape
    kt, kh, kw = taps
    rows = weight.shape[0]
    w5 = weight.float().reshape(rows, kt, kh, kw, cin).permute(0, 4, 1, 2, 3)
    xin = F.pad(x.float().permute(3, 0, 1, 2)[None], (kw // 2, kw // 2, kh // 2, kh // 2, 0, 0))
    y = F.conv3d(xin, w5)[0].permute(1, 2, 3, 0)[:T]                      # [T, H, W, rows]
    if bias is not None:
        y = y + bias.float()[:rows]
    y = y[..., :cout]
    if sub2:                                                               # stride-2 conv = odd positions
        y = y[:, 1::2, 1::2]
    if residual is not None:
        y = residual[:T].float() + y
    if out_raw is not None:
        fr = raw_frame_stride if raw_frame_stride is not None else out_raw.stride(0)
        dst = torch.as_strided(out_raw, (T, y.shape[1], y.shape[2], y.shape[3]),
                               (fr, out_raw.stride(-3), out_raw.stride(-2), 1), out_raw.storage_offset())
        dst.copy_(y.to(out_raw.dtype))
    if out_norm is not None:
        out_norm[:T].copy_(_rms(y, gamma).to(out_norm.dtype))
    if out_pix is not None:
        out_pix.copy_(y.clamp(-1, 1).permute(0, 3, 1, 2))


def vae_rmsnorm_silu(x, gamma, out, silu=True):
    out.copy_(_rms(x.float(), gamma, silu).to(out.dtype))
    return out


def vae_upsample2x(x, out):
    out.copy_(x.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2))
    return out


def softmax_rows(s, out):
    out.copy_(torch.softmax(s.float(), dim=-1).to(out.dtype))
    return out


def kv_roll(cache, dst_row, src_row, rows):
    cache[dst_row:dst_row + rows] = cache[src_row:src_row + rows].clone()


def t5_attention(q, k, v, bias_delta, key_mask, *, heads, out=None):
    """kr_t5_attn contract: no scaling, bias by offset k - q, masked keys = finfo.min, fp32 softmax."""
    L = q.shape[0]
    d = q.shape[-1] // heads
    qf = q.float().reshape(L, heads, d).transpose(0, 1)
    kf = k.float().reshape(L, heads, d).transpose(0, 1)
    vf = v.float().reshape(L, heads, d).transpose(0, 1)
    off = torch.arange(L)[None, :] - torch.arange(L)[:, None] + (L - 1)          # [q, k] -> table index
    s = qf @ kf.transpose(1, 2) + bias_delta.float()[:, off]
    if key_mask is not None:
        s = s.masked_fill(key_mask[None, None, :] == 0, torch.finfo(torch.float32).min)
    o = (torch.softmax(s, dim=-1) @ vf).transpose(0, 1).reshape(L, heads * d).to(q.dtype)
    return _store(out, o)


def linear_fp8(x, w_q, scale_w, bias=None, **kw):
    """kr_fp8_quantize + kr_gemm_fp8 contract (w_q: e4m3 bytes as uint8)."""
    amax = x.abs().max().float().clamp(min=1e-12)
    xq = (x.float() * (448.0 / amax)).clamp(-448.0, 448.0).to(torch.float8_e4m3fn).float()
    a = (xq * (amax / 448.0) * scale_w).to(torch.float32)              # fold the dequantisation into the operand
    return gemm(a, w_q.view(torch.float8_e4m3fn).float(), bias, **kw)
