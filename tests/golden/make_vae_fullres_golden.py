"""Golden for the VAE decoder at the HEADLINE geometry (latent 60x104 -> 480x832 px): the UNMODIFIED reference
``demo_utils.vae_block3.VAEDecoderWrapper`` on the CPU in fp32, block 0 (3 latent frames -> 9 frames) and one steady
block (3 -> 12 frames) of the same stream.  Pixels are stored spatially subsampled (every 4th row / column, offset 1)
as fp16 to keep the fixture small: 2 x [T, 3, 120, 208].

    python tests/golden/make_vae_fullres_golden.py     # ~10 min on 8 cores; writes tests/golden/vae_fullres.npz
"""
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(HERE.parent.parent))
import ref_shim  # noqa: E402
from realtime_video_b200.factory import synthetic_vae_params  # noqa: E402

if __name__ == "__main__":
    torch.set_num_threads(8)
    ns = ref_shim.install()
    m = ns.vae_block3.VAEDecoderWrapper()
    m.load_state_dict(synthetic_vae_params(seed=0), strict=False)
    m = m.float().eval()
    g = torch.Generator().manual_seed(31)
    out = {}
    cache = [None] * 55
    with torch.no_grad():
        for call in range(2):
            z = torch.randn(1, 3, 16, 60, 104, generator=g)
            out[f"z{call}"] = z.numpy().astype(np.float16)
            px, cache = m(z.half().float(), *cache)          # the stored fp16 latent is the input on both sides
            out[f"px{call}_sub"] = px[0, :, :, 1::4, 1::4].contiguous().numpy().astype(np.float16)
            print("call", call, tuple(px.shape), float(px.abs().mean()), flush=True)
    np.savez_compressed(HERE / "vae_fullres.npz", **out)
    print("vae_fullres.npz", sum(v.nbytes for v in out.values()) / 1e6, "MB raw")
