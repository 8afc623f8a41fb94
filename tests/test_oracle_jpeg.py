"""Pin oracle/jpeg_oracle.c to the reference's JPEG encoder: Pillow's save(format='JPEG', quality=q) — the call of
release_server.py:973 — byte for byte (integer path: bit-exact)."""
import numpy as np
import pytest

pytest.importorskip("PIL")

from oracle import jpeg_oracle
from tests.jpeg_cases import frames_fp32, images, pillow_jpeg


@pytest.mark.parametrize("hw", [(16, 16), (32, 48), (64, 96), (192, 320)])
@pytest.mark.parametrize("quality", [90, 75, 50, 100, 10])
def test_oracle_bytes_equal_pillow(hw, quality):
    for name, img in images(*hw).items():
        assert jpeg_oracle.encode_rgb8(img, quality) == pillow_jpeg(img, quality), (name, hw, quality)


def test_oracle_full_resolution_quality_90():
    """832x480 (BASELINE configs[1]) and 1280x720 (configs[3])."""
    for hw in ((480, 832), (720, 1280)):
        for name in ("noise", "smooth", "sparse"):
            img = images(*hw)[name]
            assert jpeg_oracle.encode_rgb8(img, 90) == pillow_jpeg(img, 90), (name, hw)


def test_oracle_frames_pipeline_equals_the_servers_host_path():
    """fp32 decoder frames -> normalise -> to_pil_image -> save(JPEG, 90): release_server.py:979-983 + :973."""
    import torch
    TF = pytest.importorskip("torchvision.transforms.functional")
    import io
    x = frames_fp32(3, 32, 48, seed=3)
    norm = torch.from_numpy(x.copy()).add_(1.0).mul_(0.5).clamp_(0.0, 1.0)
    want = []
    for i in range(3):
        buf = io.BytesIO()
        TF.to_pil_image(norm[i], "RGB").save(buf, format="JPEG", quality=90)
        want.append(buf.getvalue())
    assert jpeg_oracle.frames_to_jpeg(x, 90) == want


def test_oracle_header_and_coefficients():
    img = images(32, 48)["smooth"]
    data, coefs = jpeg_oracle.encode_rgb8(img, 90, return_coefs=True)
    hdr = jpeg_oracle.header(32, 48, 90)
    assert len(hdr) == 623 and data.startswith(hdr) and data.endswith(b"\xff\xd9")
    assert coefs.shape == (2 * 3 * 6, 64) and coefs.dtype == np.int16
    # decoded image is close to the source (sanity: the coefficients mean what they should)
    import io
    from PIL import Image
    dec = np.asarray(Image.open(io.BytesIO(data)).convert("RGB")).astype(np.int32)
    assert np.abs(dec - img.astype(np.int32)).mean() < 6.0


def test_oracle_rejects_unsupported_sizes():
    with pytest.raises(ValueError):
        jpeg_oracle.encode_rgb8(np.zeros((24, 32, 3), np.uint8))
