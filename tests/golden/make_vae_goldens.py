"""Generate VAE-decoder golden vectors by running the UNMODIFIED reference
``demo_utils.vae_block3.VAEDecoderWrapper`` on CPU in fp32 (build container only).

    python tests/golden/make_vae_goldens.py        # writes tests/golden/vae_small.npz

Weights are NOT stored: both sides rebuild them from ``oracle.vae_oracle.synthetic_vae_params``
(one generator per tensor, seeded by the state-dict key), so the fixture holds only the latents
and the reference's outputs.  Cases, latent 8x12 (-> 64x96 pixels) and 16x24 (-> 128x192):
  call 1: 3 latent frames with an empty cache (first-frame path: 1 + 4 + 4 = 9 frames)
  call 2: 3 latent frames with the returned cache (12 frames)
  call 3: 1 latent frame (4 frames) — single-frame chunk, exercises the `where` cache update
"""
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(HERE.parent.parent))
import ref_shim  # noqa: E402
from oracle.vae_oracle import synthetic_vae_params  # noqa: E402

OUT = {}


@torch.no_grad()
def run(tag, h, w, sub):
    ns = ref_shim.install()
    m = ns.vae_block3.VAEDecoderWrapper()
    sd = synthetic_vae_params(seed=0)
    missing = m.load_state_dict(sd, strict=False)
    assert not missing.unexpected_keys and set(missing.missing_keys) <= {"mean", "std"}, missing
    m = m.float().eval()
    g = torch.Generator().manual_seed(11)
    cache = [None] * 55
    for call, t in enumerate((3, 3, 1)):
        z = torch.randn(1, t, 16, h, w, generator=g)
        OUT[f"{tag}/z{call}"] = z.numpy()
        px, cache = m(z, *cache)
        OUT[f"{tag}/px{call}"] = px[..., ::sub, ::sub].contiguous().numpy()
        print(tag, call, tuple(px.shape), float(px.abs().mean()), flush=True)


@torch.no_grad()
def run_single(tag, h, w):
    """demo_utils.vae.VAEDecoderWrapperSingle: 3 latent frames one at a time (first flag on #0)."""
    ns = ref_shim.install()
    m = ns.vae_single.VAEDecoderWrapperSingle()
    m.load_state_dict(synthetic_vae_params(seed=0), strict=False)
    m = m.float().eval()
    shapes = [(16, h, w)] + [(384, h, w)] * 11 + [(192, 2 * h, 2 * w)] + [(384, 2 * h, 2 * w)] * 6 \
        + [(192, 4 * h, 4 * w)] * 6 + [(96, 8 * h, 8 * w)] * 7          # demo_utils/constant.py:6-39
    cache = [torch.zeros(1, c, 2, hh, ww) for (c, hh, ww) in shapes]
    g = torch.Generator().manual_seed(5)
    for i in range(3):
        z = torch.randn(1, 1, 16, h, w, generator=g)
        OUT[f"{tag}/z{i}"] = z.numpy()
        px, cache = m(z, torch.tensor(i == 0), *cache)
        OUT[f"{tag}/px{i}"] = px.contiguous().numpy()
        print(tag, i, tuple(px.shape), float(px.abs().mean()), flush=True)


@torch.no_grad()
def run_encoder(tag, H, W):
    """demo_utils.vae_block3.VAEEncoderWrapper on ONE pixel frame with an empty cache (the server's
    first-frame re-encode, release_server.py:571-576)."""
    import types
    ns = ref_shim.install()
    m = ns.vae.WanVAE_(dim=96, z_dim=16, dim_mult=[1, 2, 4, 4], num_res_blocks=2, attn_scales=[],
                       temperal_downsample=[False, True, True], dropout=0.0)
    missing = m.load_state_dict(synthetic_vae_params(seed=0, encoder=True), strict=False)
    assert not missing.unexpected_keys
    m = m.float().eval()
    enc = ns.vae_block3.VAEEncoderWrapper(types.SimpleNamespace(model=m)).float().eval()
    g = torch.Generator().manual_seed(9)
    x = torch.rand(1, 3, 1, H, W, generator=g) * 2 - 1
    mu, _ = enc(x, [None] * 55)
    OUT[f"{tag}/x"] = x.numpy()
    OUT[f"{tag}/mu"] = mu.numpy()
    print(tag, tuple(mu.shape), float(mu.abs().mean()), flush=True)


if __name__ == "__main__":
    torch.set_num_threads(8)
    run("s8x12", 8, 12, 1)
    run("s16x24", 16, 24, 2)
    run_single("single8x12", 8, 12)
    run_encoder("enc64x96", 64, 96)
    run_encoder("enc128x192", 128, 192)
    np.savez_compressed(HERE / "vae_small.npz", **OUT)
    print("vae_small.npz", sum(v.nbytes for v in OUT.values()) / 1e6, "MB raw")
