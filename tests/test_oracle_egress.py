"""The egress oracle (oracle/egress_oracle.py) against the reference's own operations
(release_server.py:979-983 + torchvision to_pil_image): bit-exact, integer output."""
import numpy as np
import pytest
import torch

from oracle.egress_oracle import frames_to_rgb8


def reference_ops(frames: torch.Tensor) -> torch.Tensor:
    """What the server does on the host: normalise in place, then to_pil_image's float branch."""
    x = frames.clone().add_(1.0).mul_(0.5).clamp_(0.0, 1.0)
    return x.mul(255).byte().movedim(-3, -1)


def edge_cases():
    vals = [-1.0, 1.0, 0.0, -0.0, -1.5, 1.5, 1 - 2 ** -24, -1 + 2 ** -24, 0.003921568859368563, 1e-8, -1e-8]
    # every fp32 boundary where (x + 1) * 0.5 * 255 crosses an integer, plus / minus one ulp
    ks = np.arange(0, 256, dtype=np.float64)
    x = (ks / 255.0 * 2.0 - 1.0).astype(np.float32)
    near = np.concatenate([x, np.nextafter(x, np.float32(2)), np.nextafter(x, np.float32(-2))])
    return np.concatenate([np.array(vals, dtype=np.float32), near])


def test_oracle_matches_reference_ops_on_random_and_boundary_values():
    g = torch.Generator().manual_seed(0)
    rnd = (torch.rand(2, 3, 37, 41, generator=g) * 2.4 - 1.2)
    assert np.array_equal(frames_to_rgb8(rnd.numpy()), reference_ops(rnd).numpy())
    e = torch.from_numpy(edge_cases())
    pad = (-e.numel()) % 3
    e = torch.cat([e, e[:pad]]).view(1, 3, 1, -1)
    assert np.array_equal(frames_to_rgb8(e.numpy()), reference_ops(e).numpy())


def test_oracle_matches_torchvision_to_pil_image():
    TF = pytest.importorskip("torchvision.transforms.functional")
    pytest.importorskip("PIL")
    g = torch.Generator().manual_seed(1)
    frames = torch.rand(1, 2, 3, 24, 40, generator=g) * 2.2 - 1.1              # [1, T, 3, H, W] like the server's tensor
    norm = frames.clone().add_(1.0).mul_(0.5).clamp_(0.0, 1.0)                  # release_server.py:983
    want = np.stack([np.asarray(TF.to_pil_image(norm[0, i], "RGB")) for i in range(2)])   # release_server.py:973
    assert np.array_equal(frames_to_rgb8(frames[0].numpy()), want)
