"""TEST INFRASTRUCTURE — CPU oracle for the causal 3D VAE decoder (not shipped, not measured).

Plain-torch restatement (NCHW, F.conv3d) of ``VAEDecoderWrapper.forward`` —
demo_utils/vae_block3.py:195-230 (wrapper), :386-443 (VAEDecoder3d.forward), :46-91 (Resample),
wan/modules/vae.py:17-36 (CausalConv3d), :39-54 (RMS_norm), :191-209 (ResidualBlock),
:229-251 (AttentionBlock).  Parameters come as a dict keyed like the reference state dict
(``decoder.conv1.weight`` ...).  The feature cache is a dict {slot: tensor [1,C,<=2,H,W]}.
Pinned by tests/golden/make_vae_goldens.py (reference run on CPU) via tests/test_oracle_vae.py.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference import this.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

MEAN = [-0.7571, -0.7089, -0.9113, 0.1075, -0.1745, 0.9653, -0.1517, 1.5508,
        0.4134, -0.0715, 0.5517, -0.3632, -0.1922, -0.9497, 0.2503, -0.2921]
STD = [2.8184, 1.4541, 2.3275, 2.6558, 1.2196, 1.7708, 2.6052, 2.0743,
       3.2687, 2.1526, 2.8652, 1.5579, 1.6382, 1.1253, 2.8251, 1.9160]


def causal_conv3d(x, w, b, cache=None):
    """vae.py:27-36: cat cache on T, zero-pad W/H by k//2 and T-front by (k_t - 1) - cache_T."""
    kt, kh, kw = w.shape[2:]
    pad_t = kt - 1
    if cache is not None and pad_t > 0:
        x = torch.cat([cache, x], dim=2)
        pad_t -= cache.shape[2]
    x = F.pad(x, (kw // 2, kw // 2, kh // 2, kh // 2, pad_t, 0))
    return F.conv3d(x, w, b)


def rms_norm(x, gamma):
    """vae.py:51-54 with channel_first: normalize over dim 1, * sqrt(C) * gamma."""
    return F.normalize(x, dim=1) * (x.shape[1] ** 0.5) * gamma


class VAEDecoderOracle:
    def __init__(self, params):
        self.p = params
        self.n_up = 15                          # decoder.upsamples.0 .. 14
        self.kinds = {3: "up3d", 7: "up3d", 11: "up2d"}

    def _cached_conv(self, name, x, cache, idx):
        """ResidualBlock / conv1 / head cache rule (vae.py:196-206): keep the last 2 input frames,
        borrowing the previous last frame when the chunk has a single frame."""
        i = idx[0]
        cx = x[:, :, -2:].clone()
        if cx.shape[2] < 2 and cache.get(i) is not None:
            cx = torch.cat([cache[i][:, :, -1:], cx], dim=2)
        y = causal_conv3d(x, self.p[name + ".weight"], self.p[name + ".bias"], cache.get(i))
        cache[i] = cx
        idx[0] += 1
        return y

    def _res(self, pre, x, cache, idx):
        p = self.p
        h = x
        if (pre + ".shortcut.weight") in p:
            h = causal_conv3d(x, p[pre + ".shortcut.weight"], p[pre + ".shortcut.bias"])
        y = F.silu(rms_norm(x, p[pre + ".residual.0.gamma"]))
        y = self._cached_conv(pre + ".residual.2", y, cache, idx)
        y = F.silu(rms_norm(y, p[pre + ".residual.3.gamma"]))
        y = self._cached_conv(pre + ".residual.6", y, cache, idx)
        return y + h

    def _attn(self, pre, x):
        p = self.p
        b, c, t, h, w = x.shape
        y = x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w)
        y = rms_norm(y, p[pre + ".norm.gamma"])
        qkv = F.conv2d(y, p[pre + ".to_qkv.weight"], p[pre + ".to_qkv.bias"])
        q, k, v = qkv.reshape(b * t, 1, 3 * c, h * w).permute(0, 1, 3, 2).chunk(3, dim=-1)
        o = F.scaled_dot_product_attention(q, k, v)
        o = o.squeeze(1).permute(0, 2, 1).reshape(b * t, c, h, w)
        o = F.conv2d(o, p[pre + ".proj.weight"], p[pre + ".proj.bias"])
        return o.reshape(b, t, c, h, w).permute(0, 2, 1, 3, 4) + x

    def _resample(self, pre, mode, x, cache, idx):
        p = self.p
        b, c, t, h, w = x.shape
        if mode == "up3d":                                   # vae_block3.py:48-66
            i = idx[0]
            if cache.get(i) is None:
                cache[i] = torch.zeros(b, c, 2, h, w, dtype=x.dtype, device=x.device)
                idx[0] += 1
            else:
                cx = x[:, :, -2:].clone()
                if cx.shape[2] < 2:
                    pad = torch.where(cache[i][:, :, -1:] == 0, torch.zeros_like(cx), cx)
                    cx = torch.cat([pad, cx], dim=2)
                y = causal_conv3d(x, p[pre + ".time_conv.weight"], p[pre + ".time_conv.bias"], cache[i])
                cache[i] = cx
                idx[0] += 1
                y = y.reshape(b, 2, c, t, h, w)
                x = torch.stack((y[:, 0], y[:, 1]), 3).reshape(b, c, t * 2, h, w)
        t = x.shape[2]
        y = x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w)
        y = F.interpolate(y.float(), scale_factor=(2.0, 2.0), mode="nearest").type_as(y)   # vae.py:57-63
        y = F.conv2d(y, p[pre + ".resample.1.weight"], p[pre + ".resample.1.bias"], padding=1)
        return y.reshape(b, t, c // 2, 2 * h, 2 * w).permute(0, 2, 1, 3, 4)

    def decode_frame(self, x, cache):
        """VAEDecoder3d.forward (vae_block3.py:386-443) on one latent frame [1, 16, 1, h, w]."""
        idx = [0]
        x = self._cached_conv("decoder.conv1", x, cache, idx)
        x = self._res("decoder.middle.0", x, cache, idx)
        x = self._attn("decoder.middle.1", x)
        x = self._res("decoder.middle.2", x, cache, idx)
        for i in range(self.n_up):
            pre = f"decoder.upsamples.{i}"
            if i in self.kinds:
                x = self._resample(pre, self.kinds[i], x, cache, idx)
            else:
                x = self._res(pre, x, cache, idx)
        x = F.silu(rms_norm(x, self.p["decoder.head.0.gamma"]))
        # head conv: same cache rule (the reference zero-fills a 2-frame buffer first, which is
        # the zero padding again, vae_block3.py:428-437)
        return self._cached_conv("decoder.head.2", x, cache, idx)

    def forward(self, z, cache):
        """z [1, T, 16, h, w] -> (pixels [1, T', 3, H, W] fp32 clamped to [-1, 1], cache)."""
        p = self.p
        z = z.permute(0, 2, 1, 3, 4)
        dt = z.dtype
        mean = torch.tensor(MEAN, dtype=torch.float32).to(dt).view(1, 16, 1, 1, 1)
        inv_std = (1.0 / torch.tensor(STD, dtype=torch.float32).to(dt)).view(1, 16, 1, 1, 1)
        z = z / inv_std + mean                                       # vae_block3.py:205-209
        x = causal_conv3d(z, p["conv2.weight"], p["conv2.bias"])
        outs = [self.decode_frame(x[:, :, i:i + 1], cache) for i in range(x.shape[2])]
        out = torch.cat(outs, dim=2).float().clamp_(-1, 1)
        return out.permute(0, 2, 1, 3, 4), cache


from realtime_video_b200.factory import synthetic_vae_params  # noqa: E402,F401  (seeded weights, shared)


class VAEEncoderOracle:
    """VAE encoder restatement.  ``encode_first``: one pixel frame, empty cache (demo_utils/vae_block3.py:136-150).
    ``forward``: the full streaming wrapper loop (vae_block3.py:141-175: chunks of 1, 4, 4, ... frames, or 4, 4,
    ... with ``stream=True`` and a warm cache) over ``encoder_chunk`` = Encoder3d.forward with the feature cache
    (wan/modules/vae.py:301-345; cached causal convs :17-36 / :191-206; Resample downsample2d/3d :144-172 — the
    downsample3d time_conv (3,1,1)/stride (2,1,1) is skipped while its cache slot is None, afterwards it sees
    [last cached frame, chunk] and the slot keeps the chunk's last frame), then WanVAE_.conv1, chunk(2)[0] and
    (mu - mean) * 1/std (vae_block3.py:166-172).  Params keyed 'encoder.*', 'conv1.*'; cache = dict slot -> tensor.
    Pinned by tests/golden/make_vae_goldens.py and make_vae_encoder_stream_goldens.py (reference run on CPU)."""

    def __init__(self, params):
        self.p = params
        self.dec = VAEDecoderOracle(params)     # reuses causal-conv / res / attention restatements

    def encoder_chunk(self, x, cache):
        """Encoder3d.forward(x [1, 3, t, H, W], feat_cache) -> [1, 32, t', H/8, W/8]."""
        p, d = self.p, self.dec
        idx = [0]
        x = d._cached_conv("encoder.conv1", x, cache, idx)
        n = 0
        for i in range(4):
            for _ in range(2):
                x = d._res(f"encoder.downsamples.{n}", x, cache, idx)
                n += 1
            if i != 3:
                pre = f"encoder.downsamples.{n}"
                b, c, t, h, w = x.shape
                y = x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w)
                y = F.conv2d(F.pad(y, (0, 1, 0, 1)), p[pre + ".resample.1.weight"], p[pre + ".resample.1.bias"], stride=2)
                x = y.reshape(b, t, c, h // 2, w // 2).permute(0, 2, 1, 3, 4)
                if i in (1, 2):                                          # downsample3d (vae.py:157-172)
                    j = idx[0]
                    if cache.get(j) is None:
                        cache[j] = x.clone()
                    else:
                        last = x[:, :, -1:].clone()
                        x = F.conv3d(torch.cat([cache[j][:, :, -1:], x], dim=2), p[pre + ".time_conv.weight"],
                                     p[pre + ".time_conv.bias"], stride=(2, 1, 1))
                        cache[j] = last
                    idx[0] += 1
                n += 1
        x = d._res("encoder.middle.0", x, cache, idx)
        x = d._attn("encoder.middle.1", x)
        x = d._res("encoder.middle.2", x, cache, idx)
        x = F.silu(rms_norm(x, p["encoder.head.0.gamma"]))
        return d._cached_conv("encoder.head.2", x, cache, idx)

    def forward(self, z, cache, stream=False):
        """VAEEncoderWrapper.forward: z [1, 3, T, H, W] pixels, cache dict -> (mu [1, 16, T', H/8, W/8], cache)."""
        t = z.shape[2]
        out = None
        offset = 1
        for i in range(1 + (t - 1) // 4):
            if i == 0 and cache.get(0) is None:
                out = self.encoder_chunk(z[:, :, :1], cache)
            else:
                start = i - 1
                if stream:
                    offset, start = 0, i
                o = self.encoder_chunk(z[:, :, offset + 4 * start:offset + 4 * (start + 1)], cache)
                out = o if (i == 0 and stream) else torch.cat([out, o], dim=2)
        mu = causal_conv3d(out, self.p["conv1.weight"], self.p["conv1.bias"]).chunk(2, dim=1)[0]
        dt = mu.dtype
        mean = torch.tensor(MEAN, dtype=torch.float32).to(dt).view(1, 16, 1, 1, 1)
        inv_std = (1.0 / torch.tensor(STD, dtype=torch.float32).to(dt)).view(1, 16, 1, 1, 1)
        return (mu - mean) * inv_std, cache

    def encode_first(self, x):
        """x [1, 3, 1, H, W] -> mu [1, 16, 1, H/8, W/8]."""
        p, d = self.p, self.dec
        cache, idx = {}, [0]
        x = d._cached_conv("encoder.conv1", x, cache, idx)
        n = 0
        for i in range(4):
            for _ in range(2):
                x = d._res(f"encoder.downsamples.{n}", x, cache, idx)
                n += 1
            if i != 3:
                pre = f"encoder.downsamples.{n}"
                b, c, t, h, w = x.shape
                y = x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w)
                y = F.conv2d(F.pad(y, (0, 1, 0, 1)), p[pre + ".resample.1.weight"], p[pre + ".resample.1.bias"], stride=2)
                x = y.reshape(b, t, c, h // 2, w // 2).permute(0, 2, 1, 3, 4)
                n += 1          # first chunk: downsample3d only stores its cache (vae.py:160-163)
        x = d._res("encoder.middle.0", x, cache, idx)
        x = d._attn("encoder.middle.1", x)
        x = d._res("encoder.middle.2", x, cache, idx)
        x = F.silu(rms_norm(x, p["encoder.head.0.gamma"]))
        x = d._cached_conv("encoder.head.2", x, cache, idx)
        mu = causal_conv3d(x, p["conv1.weight"], p["conv1.bias"]).chunk(2, dim=1)[0]
        dt = mu.dtype
        mean = torch.tensor(MEAN, dtype=torch.float32).to(dt).view(1, 16, 1, 1, 1)
        inv_std = (1.0 / torch.tensor(STD, dtype=torch.float32).to(dt)).view(1, 16, 1, 1, 1)
        return (mu - mean) * inv_std
