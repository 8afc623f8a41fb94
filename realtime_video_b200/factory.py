"""Build the hot-path models with seeded synthetic weights (there are no checkpoints on the
build or GPU boxes): the reference's init (xavier / N(0,.02)) with
``head.head.weight`` redrawn N(0,.02) because the reference zero-initialises it
(causal_model.py:1173) and a zero head makes every output identically zero."""
from __future__ import annotations

import torch

from .wan_wrapper import KNOWN_CONFIGS, WanDiffusionWrapper


def synthetic_transformer(size: str = "14B", device="cuda", dtype=torch.bfloat16, seed: int = 0,
                          timestep_shift: float = 5.0, num_layers: int | None = None,
                          **overrides) -> WanDiffusionWrapper:
    cfg = dict(KNOWN_CONFIGS[size]) if size in KNOWN_CONFIGS else {}
    cfg.update(overrides)
    if num_layers is not None:
        cfg["num_layers"] = num_layers
    torch.manual_seed(seed)
    if torch.device(device).type == "cuda":
        torch.cuda.manual_seed(seed)
    w = WanDiffusionWrapper(model_name="synthetic-" + size, timestep_shift=timestep_shift, is_causal=True,
                            model_config=cfg, device=device, dtype=dtype)
    with torch.no_grad():
        w.model.head.head.weight.normal_(std=0.02)
    w.eval().requires_grad_(False)
    for blk in w.model.blocks:          # release_server.py:176-177
        blk.self_attn.fuse_projections()
    return w


def synthetic_prompt_embeds(device="cuda", text_dim: int = 4096, tokens: int = 64, seed: int = 1):
    """randn [1, 512, text_dim] with rows >= tokens zeroed (utils/wan_wrapper.py:52-53)."""
    g = torch.Generator().manual_seed(seed)
    e = torch.randn(1, 512, text_dim, generator=g)
    e[:, tokens:] = 0
    return e.to(device=device, dtype=torch.bfloat16)


def synthetic_vae_params(seed: int = 0, dim: int = 96, z_dim: int = 16, encoder: bool = False):
    """Deterministic decoder weights independent of module construction order: every tensor is
    drawn from its own generator seeded by (seed, key).  Conv weights ~ N(0, 1/fan_in) * 1.4,
    biases ~ N(0, .02), gammas ~ 1 + N(0, .1)."""
    import zlib
    shapes = {}
    dims = [dim * 4, dim * 4, dim * 4, dim * 2, dim]

    def conv3(name, ci, co, k):
        shapes[name + ".weight"] = (co, ci, *k)
        shapes[name + ".bias"] = (co,)

    def res(pre, ci, co):
        shapes[pre + ".residual.0.gamma"] = (ci, 1, 1, 1)
        conv3(pre + ".residual.2", ci, co, (3, 3, 3))
        shapes[pre + ".residual.3.gamma"] = (co, 1, 1, 1)
        conv3(pre + ".residual.6", co, co, (3, 3, 3))
        if ci != co:
            conv3(pre + ".shortcut", ci, co, (1, 1, 1))

    conv3("conv2", z_dim, z_dim, (1, 1, 1))
    conv3("decoder.conv1", z_dim, dims[0], (3, 3, 3))
    res("decoder.middle.0", dims[0], dims[0])
    shapes["decoder.middle.1.norm.gamma"] = (dims[0], 1, 1)
    shapes["decoder.middle.1.to_qkv.weight"] = (3 * dims[0], dims[0], 1, 1)
    shapes["decoder.middle.1.to_qkv.bias"] = (3 * dims[0],)
    shapes["decoder.middle.1.proj.weight"] = (dims[0], dims[0], 1, 1)
    shapes["decoder.middle.1.proj.bias"] = (dims[0],)
    res("decoder.middle.2", dims[0], dims[0])
    n = 0
    cin = dims[0]
    plan = [(dims[0], dims[1], "up3d"), (dims[1] // 2, dims[2], "up3d"), (dims[2] // 2, dims[3], "up2d"),
            (dims[3] // 2, dims[4], None)]
    for ci, co, up in plan:
        cin = ci
        for _ in range(3):
            res(f"decoder.upsamples.{n}", cin, co)
            cin = co
            n += 1
        if up is not None:
            pre = f"decoder.upsamples.{n}"
            shapes[pre + ".resample.1.weight"] = (co // 2, co, 3, 3)
            shapes[pre + ".resample.1.bias"] = (co // 2,)
            if up == "up3d":
                conv3(pre + ".time_conv", co, 2 * co, (3, 1, 1))
            n += 1
    shapes["decoder.head.0.gamma"] = (dims[4], 1, 1, 1)
    conv3("decoder.head.2", dims[4], 3, (3, 3, 3))
    if encoder:
        # Encoder3d (wan/modules/vae.py:254-299) + WanVAE_.conv1 (:471): keys 'encoder.*', 'conv1.*'
        shapes = {k: v for k, v in shapes.items() if False}
        edims = [dim, dim, dim * 2, dim * 4, dim * 4]
        conv3("conv1", 2 * z_dim, 2 * z_dim, (1, 1, 1))
        conv3("encoder.conv1", 3, edims[0], (3, 3, 3))
        n = 0
        for i, (ci, co) in enumerate(zip(edims[:-1], edims[1:])):
            for _ in range(2):
                res(f"encoder.downsamples.{n}", ci, co)
                ci = co
                n += 1
            if i != 3:
                pre = f"encoder.downsamples.{n}"
                shapes[pre + ".resample.1.weight"] = (co, co, 3, 3)
                shapes[pre + ".resample.1.bias"] = (co,)
                if i >= 1:                      # temperal_downsample = [False, True, True]
                    conv3(pre + ".time_conv", co, co, (3, 1, 1))
                n += 1
        top = edims[-1]
        res("encoder.middle.0", top, top)
        shapes["encoder.middle.1.norm.gamma"] = (top, 1, 1)
        shapes["encoder.middle.1.to_qkv.weight"] = (3 * top, top, 1, 1)
        shapes["encoder.middle.1.to_qkv.bias"] = (3 * top,)
        shapes["encoder.middle.1.proj.weight"] = (top, top, 1, 1)
        shapes["encoder.middle.1.proj.bias"] = (top,)
        res("encoder.middle.2", top, top)
        shapes["encoder.head.0.gamma"] = (top, 1, 1, 1)
        conv3("encoder.head.2", top, 2 * z_dim, (3, 3, 3))
    out = {}
    for k, shp in shapes.items():
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(k.encode())) % (2 ** 31))
        t = torch.randn(shp, generator=g)
        if k.endswith("gamma"):
            t = 1.0 + 0.1 * t
        elif k.endswith("bias"):
            t = 0.02 * t
        else:
            fan_in = 1
            for d in shp[1:]:
                fan_in *= d
            t = t * (1.4 / fan_in ** 0.5)
        out[k] = t
    return out


def synthetic_vae_decoder(device="cuda", dtype=torch.float16, seed: int = 0):
    """VAEDecoderWrapper (demo_utils/vae_block3.py:177) with seeded synthetic weights."""
    from .vae import VAEDecoderWrapper
    m = VAEDecoderWrapper()
    m.load_state_dict(synthetic_vae_params(seed=seed), strict=False)
    return m.to(device=device, dtype=dtype).eval().requires_grad_(False)


def synthetic_vae_encoder(device="cuda", dtype=torch.float16, seed: int = 0):
    """VAEEncoderWrapper (demo_utils/vae_block3.py:116) with seeded synthetic weights."""
    from .vae import VAEEncoderWrapper
    m = VAEEncoderWrapper()
    m.load_state_dict(synthetic_vae_params(seed=seed, encoder=True), strict=False)
    return m.to(device=device, dtype=dtype).eval().requires_grad_(False)
