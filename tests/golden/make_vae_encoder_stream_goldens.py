"""Golden vectors for the STREAMING VAE encode from the UNMODIFIED reference on CPU (fp32):
demo_utils.vae_block3.VAEEncoderWrapper.forward(z, feat_cache, stream) over wan.modules.vae.Encoder3d with the
synthetic weights of realtime_video_b200.factory.synthetic_vae_params(seed=0, encoder=True).

Cases (the two ways release_server.py calls the encoder, :518-525 webcam, :531-538 / :585 v2v / start frame):
  cold9_64x96    x [1,3,9,64,96], cache [None]*55, stream=False      -> mu [1,16,3,8,12]   (1 + 4 + 4 frames)
  stream8_64x96  next 8 frames, the cache returned above, stream=True -> mu [1,16,2,8,12]   (4 + 4 frames)
  stream4_64x96  next 4 frames, stream=True                           -> mu [1,16,1,8,12]
  cold5_48x80    x [1,3,5,48,80], cold, stream=False                  -> mu [1,16,2,6,10]   (ragged grid)
Run in the build container only:  python tests/golden/make_vae_encoder_stream_goldens.py
"""
import sys
import types
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
sys.path.insert(0, str(HERE))
import ref_shim  # noqa: E402
from realtime_video_b200.factory import synthetic_vae_params  # noqa: E402

OUT = {}


def encoder():
    ns = ref_shim.install()
    m = ns.vae.WanVAE_(dim=96, z_dim=16, dim_mult=[1, 2, 4, 4], num_res_blocks=2, attn_scales=[],
                       temperal_downsample=[False, True, True], dropout=0.0)
    missing = m.load_state_dict(synthetic_vae_params(seed=0, encoder=True), strict=False)
    assert not missing.unexpected_keys
    m = m.float().eval()
    return ns.vae_block3.VAEEncoderWrapper(types.SimpleNamespace(model=m)).float().eval()


@torch.no_grad()
def main():
    torch.set_num_threads(8)
    enc = encoder()
    g = torch.Generator().manual_seed(21)
    x = torch.rand(1, 3, 21, 64, 96, generator=g) * 2 - 1
    # smooth the clip a little in time so consecutive frames are correlated like video
    x = (x + torch.roll(x, 1, dims=2) + torch.roll(x, 2, dims=2)) / 3
    cache = [None] * 55
    mu, cache = enc(x[:, :, :9], cache, stream=False)
    OUT["cold9_64x96/x"], OUT["cold9_64x96/mu"] = x[:, :, :9].numpy(), mu.numpy()
    mu, cache = enc(x[:, :, 9:17], cache, stream=True)
    OUT["stream8_64x96/x"], OUT["stream8_64x96/mu"] = x[:, :, 9:17].numpy(), mu.numpy()
    mu, cache = enc(x[:, :, 17:21], cache, stream=True)
    OUT["stream4_64x96/x"], OUT["stream4_64x96/mu"] = x[:, :, 17:21].numpy(), mu.numpy()
    n_cached = sum(c is not None for c in cache)
    OUT["n_cache_slots"] = np.array([n_cached])
    y = torch.rand(1, 3, 5, 48, 80, generator=g) * 2 - 1
    mu, _ = enc(y, [None] * 55, stream=False)
    OUT["cold5_48x80/x"], OUT["cold5_48x80/mu"] = y.numpy(), mu.numpy()
    for k, v in OUT.items():
        print(k, v.shape, float(np.abs(v).mean()))
    np.savez_compressed(HERE / "vae_encoder_stream.npz", **OUT)
    print("vae_encoder_stream.npz", sum(v.nbytes for v in OUT.values()) / 1e6, "MB raw")


if __name__ == "__main__":
    main()
