"""The device JPEG encoder's per-thread code (realtime_video_b200/csrc/kr_jpeg_core.cuh) compiled for the HOST and run
with loops in place of the CUDA grid (tests/jpeg_emulate.cpp), against Pillow and the oracle: every byte equal.
Covers both loaders (fp32 decoder frames, RGB bytes), several frames per call, different CTA sizes for the per-frame
passes (chunking of the prefix sums), a garbage-initialised workspace and the capacity-overflow contract."""
import ctypes
import shutil
import subprocess
from pathlib import Path

import numpy as np
import pytest

pytest.importorskip("PIL")
pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not on PATH")

from oracle import jpeg_oracle
from tests.jpeg_cases import frames_fp32, images, pillow_jpeg

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    so = tmp_path_factory.mktemp("jpeg_emulate") / "libjpeg_emulate.so"
    subprocess.run(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-x", "c++",
                    str(ROOT / "tests" / "jpeg_emulate.cpp"), "-o", str(so)], check=True)
    lib = ctypes.CDLL(str(so))
    lib.jpeg_emulate.restype = ctypes.c_int
    lib.jpeg_emulate.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                 ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    return lib


def run(lib, src: np.ndarray, kind: int, quality=90, threads=1024, cap=None, want_coefs=False):
    if kind == 0:
        T, _, H, W = src.shape
    else:
        T, H, W, _ = src.shape
    cap = cap or (H * W * 3 + 4096) // 4 * 4
    out = np.full((T, cap + 64), 0xEE, np.uint8)               # 64 guard bytes per frame
    sizes = np.zeros(T, np.int32)
    coefs = np.zeros((T, H // 16 * (W // 16) * 6, 64), np.int16) if want_coefs else None
    # frames are laid out with pitch cap: emulate on a compact buffer, then check the guard separately
    buf = np.full((T, cap), 0xEE, np.uint8)
    rc = lib.jpeg_emulate(src.ctypes.data, kind, T, H, W, quality, buf.ctypes.data, cap, sizes.ctypes.data, threads,
                          coefs.ctypes.data if want_coefs else None)
    assert rc == 0, rc
    files = [buf[t, :sizes[t]].tobytes() if sizes[t] > 0 else None for t in range(T)]
    return files, sizes, buf, coefs


@pytest.mark.parametrize("hw", [(16, 16), (32, 48), (64, 96)])
@pytest.mark.parametrize("quality", [90, 75, 100, 10])
@pytest.mark.parametrize("threads", [1024, 64, 7])
def test_emulated_kernels_equal_pillow_rgb8(emu, hw, quality, threads):
    imgs = images(*hw)
    frames = np.ascontiguousarray(np.stack(list(imgs.values())))
    files, sizes, _, _ = run(emu, frames, 1, quality, threads)
    for (name, img), got in zip(imgs.items(), files):
        assert got == pillow_jpeg(img, quality), (name, hw, quality, threads)


def test_emulated_kernels_full_resolution(emu):
    imgs = images(480, 832)
    pick = ["noise", "smooth", "sparse", "black"]
    frames = np.ascontiguousarray(np.stack([imgs[k] for k in pick]))
    files, _, _, coefs = run(emu, frames, 1, 90, 1024, want_coefs=True)
    for i, k in enumerate(pick):
        want, want_coefs = jpeg_oracle.encode_rgb8(imgs[k], 90, return_coefs=True)
        assert np.array_equal(coefs[i], want_coefs), k
        assert files[i] == want == pillow_jpeg(imgs[k], 90), k


def test_emulated_fp32_loader_equals_the_servers_host_path(emu):
    x = frames_fp32(3, 64, 96, seed=5)
    files, _, _, _ = run(emu, x, 0, 90, 1024)
    assert files == jpeg_oracle.frames_to_jpeg(x, 90)


def test_capacity_overflow_is_reported_and_nothing_is_written_past_cap(emu):
    img = images(32, 48)["noise"]
    need = len(pillow_jpeg(img, 90))
    frames = np.ascontiguousarray(np.stack([img, images(32, 48)["black"]]))
    cap = (need - 40) // 4 * 4                                # too small for the noise frame, enough for the flat one
    T = 2
    big = np.full(T * cap + 256, 0xEE, np.uint8)
    sizes = np.zeros(T, np.int32)
    rc = emu.jpeg_emulate(frames.ctypes.data, 1, T, 32, 48, 90, big.ctypes.data, cap, sizes.ctypes.data, 128, None)
    assert rc == 0
    assert sizes[0] == -need
    assert sizes[1] == len(pillow_jpeg(frames[1], 90))
    assert big[cap:cap + sizes[1]].tobytes() == pillow_jpeg(frames[1], 90)     # frame 1 intact, right after cap
    assert (big[T * cap:] == 0xEE).all()
    assert big[:cap].tobytes() == pillow_jpeg(img, 90)[:cap]                   # the truncated prefix is still right


def test_python_op_argument_plumbing_with_the_emulator_behind_the_abi(emu, monkeypatch):
    """ops.frames_to_jpeg / ops.jpeg_files on CPU tensors with the C ABI replaced by the host emulator: checks the
    wrapper's shape handling, default capacity, argument order and the host-side file extraction."""
    import torch
    from realtime_video_b200 import _lib, ops

    real = _lib.load()

    class FakeLib:
        def kr_jpeg_workspace_bytes(self, T, H, W):
            return real.kr_jpeg_workspace_bytes(T, H, W)               # host-only size query of the real library

        def _run(self, kind, src, T, H, W, q, out, cap, sizes, ws, ws_bytes, stream):
            assert ws_bytes >= self.kr_jpeg_workspace_bytes(T, H, W) > 0
            return emu.jpeg_emulate(src, kind, T, H, W, q, out, cap, sizes, 1024, None)

        def kr_frames_to_jpeg(self, *a):
            return self._run(0, *a)

        def kr_rgb8_to_jpeg(self, *a):
            return self._run(1, *a)

    monkeypatch.setattr(ops._lib, "load", lambda: FakeLib())
    monkeypatch.setattr(ops._lib, "check", lambda rc, what: None if rc == 0 else (_ for _ in ()).throw(AssertionError(rc)))
    monkeypatch.setattr(ops, "_req", lambda t, name, dtype=None: None)
    monkeypatch.setattr(ops, "_stream", lambda: 0)
    x = frames_fp32(4, 32, 48, seed=9)
    out, sizes = ops.frames_to_jpeg(torch.from_numpy(x)[None], 90)            # [1, T, 3, H, W] like the server's tensor
    assert tuple(out.shape) == (4, (32 * 48 * 3 + 4096 + 3) // 4 * 4) and sizes.dtype == torch.int32
    assert ops.jpeg_files(out, sizes) == jpeg_oracle.frames_to_jpeg(x, 90)
    img = images(32, 48)["smooth"]
    out, sizes = ops.frames_to_jpeg(torch.from_numpy(img)[None], 75)
    assert ops.jpeg_files(out, sizes) == [pillow_jpeg(img, 75)]


def test_random_images_sizes_and_qualities_property(emu):
    """Property test: for random image content (noise mixed with flat regions and saturated pixels), random sizes
    (multiples of 16) and every quality from 1 to 100, the emulated device encoder, the oracle and Pillow agree on
    every byte."""
    hyp = pytest.importorskip("hypothesis")
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(h=st.integers(1, 4), w=st.integers(1, 5), quality=st.integers(1, 100), seed=st.integers(0, 2 ** 31 - 1),
           kind=st.sampled_from(["noise", "flat+noise", "saturated", "lines"]), threads=st.sampled_from([1024, 96, 5]))
    def check(h, w, quality, seed, kind, threads):
        H, W = 16 * h, 16 * w
        rng = np.random.default_rng(seed)
        img = (rng.random((H, W, 3)) * 256).astype(np.uint8)
        if kind == "flat+noise":
            img[: H // 2] = rng.integers(0, 256, 3, dtype=np.uint8)
        elif kind == "saturated":
            img = np.where(rng.random((H, W, 3)) < 0.5, 0, 255).astype(np.uint8)
        elif kind == "lines":
            img[:] = 128
            img[::3] = 255
            img[:, ::5] = 0
        img = np.ascontiguousarray(img)
        files, _, _, _ = run(emu, img[None], 1, quality, threads)
        want = pillow_jpeg(img, quality)
        assert files[0] == want and jpeg_oracle.encode_rgb8(img, quality) == want

    check()
