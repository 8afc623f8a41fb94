"""GPU parity of the device JPEG encoder (kr_frames_to_jpeg / kr_rgb8_to_jpeg through the C ABI) with the reference's
encoder: every file byte-identical to Pillow's save(format='JPEG', quality=q) (release_server.py:973) and to the
oracle.  Integer / byte path: bit-exact.  (Runs last: the newest kernels of the suite.)"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from tests.jpeg_cases import frames_fp32, images, pillow_jpeg


def files_of(out, sizes):
    from realtime_video_b200 import ops
    return ops.jpeg_files(out, sizes)


@pytest.mark.parametrize("hw", [(16, 16), (32, 48), (64, 96), (192, 320)])
@pytest.mark.parametrize("quality", [90, 75, 100, 10])
def test_rgb8_to_jpeg_equals_pillow(hw, quality):
    from realtime_video_b200 import ops
    imgs = images(*hw)
    frames = torch.from_numpy(np.stack(list(imgs.values()))).cuda()
    out, sizes = ops.frames_to_jpeg(frames, quality)
    got = files_of(out, sizes)
    for (name, img), g in zip(imgs.items(), got):
        assert g == pillow_jpeg(img, quality), (name, hw, quality)


@pytest.mark.parametrize("hw", [(480, 832), (720, 1280)])
def test_full_resolution_block_of_12_frames(hw):
    """BASELINE configs[1] / configs[3] frame sizes, 12 frames per call like one block of the server."""
    from oracle import jpeg_oracle
    from realtime_video_b200 import ops
    imgs = images(*hw)
    names = list(imgs)
    frames_np = np.stack([imgs[names[i % len(names)]] for i in range(12)])
    out, sizes = ops.frames_to_jpeg(torch.from_numpy(frames_np).cuda(), 90)
    got = files_of(out, sizes)
    for i in range(12):
        want = jpeg_oracle.encode_rgb8(frames_np[i], 90)
        assert got[i] == want, (names[i % len(names)], hw)
        if i < len(names):
            assert want == pillow_jpeg(frames_np[i], 90)
    # calling again on the same stream reuses the workspace: same bytes
    out2, sizes2 = ops.frames_to_jpeg(torch.from_numpy(frames_np).cuda(), 90)
    assert files_of(out2, sizes2) == got


def test_fp32_decoder_frames_equal_the_servers_host_path():
    """fp32 [1, T, 3, H, W] -> normalise -> to_pil_image -> save(JPEG, 90) on the host == one device call."""
    from oracle import jpeg_oracle
    from realtime_video_b200 import ops
    x = frames_fp32(12, 96, 160, seed=11)
    out, sizes = ops.frames_to_jpeg(torch.from_numpy(x).cuda()[None], 90)
    assert files_of(out, sizes) == jpeg_oracle.frames_to_jpeg(x, 90)
    # and through the byte kernel: frames_to_rgb8 -> rgb8_to_jpeg gives the same files
    rgb = ops.frames_to_rgb8(torch.from_numpy(x).cuda())
    out2, sizes2 = ops.frames_to_jpeg(rgb, 90)
    assert files_of(out2, sizes2) == files_of(out, sizes)


def test_decoder_to_jpeg_end_to_end():
    """VAE decode -> kr_frames_to_jpeg == VAE decode -> the reference's host-side normalise + Pillow."""
    import io
    TF = pytest.importorskip("torchvision.transforms.functional")
    from realtime_video_b200 import ops
    from realtime_video_b200.factory import synthetic_vae_decoder
    vae = synthetic_vae_decoder(device="cuda")
    z = torch.randn(1, 3, 16, 8, 12, device="cuda").half()
    px, _ = vae(z, *([None] * 55))                                        # [1, 9, 3, 64, 96] fp32
    norm = px.cpu().add_(1.0).mul_(0.5).clamp_(0.0, 1.0)                  # release_server.py:983
    want = []
    for i in range(px.shape[1]):
        buf = io.BytesIO()
        TF.to_pil_image(norm[0, i], "RGB").save(buf, format="JPEG", quality=90)      # release_server.py:973
        want.append(buf.getvalue())
    out, sizes = ops.frames_to_jpeg(px, 90)
    assert files_of(out, sizes) == want


def test_capacity_overflow_and_argument_checks():
    from realtime_video_b200 import _lib, ops
    img = images(32, 48)["noise"]
    need = len(pillow_jpeg(img, 90))
    frames = torch.from_numpy(np.stack([img, images(32, 48)["black"]])).cuda()
    cap = (need - 40) // 4 * 4
    guard = torch.full((2 * cap + 256,), 0xEE, dtype=torch.uint8, device="cuda")
    out, sizes = ops.frames_to_jpeg(frames, 90, cap=cap, out=guard[:2 * cap].view(2, cap))
    s = sizes.cpu().tolist()
    assert s[0] == -need and s[1] == len(pillow_jpeg(frames[1].cpu().numpy(), 90))
    h = guard.cpu().numpy()
    assert (h[2 * cap:] == 0xEE).all()                                    # nothing written past the buffers
    assert h[:cap].tobytes() == pillow_jpeg(img, 90)[:cap]
    assert h[cap:cap + s[1]].tobytes() == pillow_jpeg(frames[1].cpu().numpy(), 90)
    with pytest.raises(_lib.KreaB200Error):
        ops.jpeg_files(out, sizes)
    with pytest.raises(_lib.KreaB200Error):
        ops.frames_to_jpeg(torch.zeros(1, 3, 24, 32, device="cuda"))      # 24 is not a multiple of 16
    with pytest.raises(_lib.KreaB200Error):
        ops.frames_to_jpeg(torch.zeros(1, 3, 16, 16))                     # CPU tensor: no fallback
