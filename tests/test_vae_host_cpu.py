"""Host logic of realtime_video_b200/vae.py (DecoderEngine / EncoderEngine) executed on the CPU with the
kernels replaced by the fp32 stand-ins of tests/cpu_ops_emulation.py: persistent conv-input buffers whose two
leading frames ARE the feature cache, the cache roll, the first-frame sentinel and the ``where`` quirk of the
up3d time_conv cache, the channel->time interleave, chunking, cache export/import, the single-frame wrapper and
the first-chunk encoder — against the goldens of the reference's own VAE modules (tests/golden/vae_small.npz).
Activations are stored in fp16 like the product (the engine picks the dtype); tolerance rel-L2 <= 5e-3."""
import pytest
import torch

from oracle.vae_oracle import synthetic_vae_params
from tests import cpu_ops_emulation as emu
from tests.golden_io import load_npz, rel_l2

TOL = 5e-3


@pytest.fixture(autouse=True)
def cpu_ops(monkeypatch):
    import realtime_video_b200.vae as vae
    monkeypatch.setattr(vae, "ops", emu)


@pytest.fixture(scope="module")
def g():
    return load_npz("vae_small.npz")


def decoder():
    from realtime_video_b200.vae import VAEDecoderWrapper
    m = VAEDecoderWrapper()
    m.load_state_dict(synthetic_vae_params(seed=0), strict=False)
    return m.half().eval()


def test_streaming_decode_three_calls(g):
    m = decoder()
    cache = [None] * 55
    with torch.no_grad():
        for call, frames in enumerate([9, 12, 4]):
            px, cache = m(g[f"s8x12/z{call}"].half(), *cache)
            assert px.dtype == torch.float32 and px.shape[1] == frames
            r = rel_l2(px, g[f"s8x12/px{call}"])
            assert r < TOL, f"call {call}: rel_l2={r:.3e}"


def test_reset_and_cache_roundtrip(g):
    m = decoder()
    z0, z1 = g["s8x12/z0"].half(), g["s8x12/z1"].half()
    with torch.no_grad():
        a, cache = m(z0, *([None] * 55))
        b, _ = m(z1, *cache)
        a2, cache2 = m(z0, *([None] * 55))
        b2, _ = m(z1, *[c.clone() if c is not None else None for c in cache2])
    assert torch.equal(a, a2) and torch.equal(b, b2)
    with pytest.raises(ValueError, match="partial VAE feature cache"):
        m(z1, *([cache2[0]] + [None] * 54))


def test_single_frame_wrapper(g):
    from realtime_video_b200.vae import VAEDecoderWrapperSingle
    m = VAEDecoderWrapperSingle()
    m.load_state_dict(synthetic_vae_params(seed=0), strict=False)
    m = m.half().eval()
    h, w = 8, 12
    shapes = [(16, h, w)] + [(384, h, w)] * 11 + [(192, 2 * h, 2 * w)] + [(384, 2 * h, 2 * w)] * 6 \
        + [(192, 4 * h, 4 * w)] * 6 + [(96, 8 * h, 8 * w)] * 7
    cache = [torch.zeros(1, c, 2, hh, ww, dtype=torch.float16) for (c, hh, ww) in shapes]
    with torch.no_grad():
        for i in range(3):
            px, cache = m(g[f"single8x12/z{i}"].half(), torch.tensor(i == 0), *cache)
            assert px.shape == (1, 4, 3, 64, 96) and len(cache) == 32
            r = rel_l2(px.float(), g[f"single8x12/px{i}"])
            assert r < TOL, f"single frame {i}: rel_l2={r:.3e}"


def test_encoder_first_frame(g):
    from realtime_video_b200.vae import VAEEncoderWrapper
    m = VAEEncoderWrapper()
    m.load_state_dict(synthetic_vae_params(seed=0, encoder=True), strict=False)
    m = m.half().eval()
    with torch.no_grad():
        mu, _ = m(g["enc64x96/x"].half(), [None] * 55)
    assert mu.shape == g["enc64x96/mu"].shape
    r = rel_l2(mu, g["enc64x96/mu"])
    assert r < TOL, f"rel_l2={r:.3e}"


def test_streaming_encoder_chunks(g):
    """1 + 4 + 4 frames on a fresh cache, then 4 + 4 and 4 frames with stream=True on the carried cache — the two
    ways the server calls the encoder (release_server.py:518-538) — against the reference's VAEEncoderWrapper."""
    from realtime_video_b200.vae import VAEEncoderWrapper
    gs = load_npz("vae_encoder_stream.npz")
    m = VAEEncoderWrapper()
    m.load_state_dict(synthetic_vae_params(seed=0, encoder=True), strict=False)
    m = m.half().eval()
    cache = [None] * 55
    with torch.no_grad():
        for tag, stream in (("cold9_64x96", False), ("stream8_64x96", True), ("stream4_64x96", True)):
            mu, cache = m(gs[f"{tag}/x"].half(), cache, stream=stream)
            assert mu.shape == gs[f"{tag}/mu"].shape and len(cache) == 55
            r = rel_l2(mu, gs[f"{tag}/mu"])
            assert r < TOL, f"{tag}: rel_l2={r:.3e}"
        assert sum(c is not None for c in cache) == int(gs["n_cache_slots"][0])
        # a fresh cache restarts the stream: same first latent frame as before
        again, _ = m(gs["cold9_64x96/x"][:, :, :1].half(), [None] * 55)
        assert rel_l2(again, gs["cold9_64x96/mu"][:, :, :1]) < TOL
        # 1 + 4k + r frames: the trailing r < 4 frames are ignored like in the reference loop (1 + (t - 1) // 4 chunks)
        seven, cache7 = m(gs["cold9_64x96/x"][:, :, :7].half(), [None] * 55)
        assert seven.shape[2] == 2 and rel_l2(seven, gs["cold9_64x96/mu"][:, :, :2]) < TOL
        with pytest.raises(ValueError, match="needs 4"):          # a ragged streaming chunk cannot be encoded
            m(gs["stream8_64x96/x"][:, :, :6].half(), cache7, stream=True)


def test_classic_wrapper_encode_to_latent(g):
    """WanVAEWrapper.encode_to_latent (utils/wan_wrapper.py:80-96): per sample a fresh 1 + 4 + 4 stream; the
    module tree accepts the reference's full WanVAE_ state dict (encoder.*, conv1.*, conv2.*, decoder.*)."""
    import realtime_video_b200.wan_wrapper as ww
    gs = load_npz("vae_encoder_stream.npz")
    w = ww.WanVAEWrapper(load_pretrained=False)
    sd = dict(synthetic_vae_params(seed=0, encoder=True))
    sd.update(synthetic_vae_params(seed=0))
    missing = w.model.load_state_dict(sd, strict=False)
    assert not missing.unexpected_keys and set(missing.missing_keys) <= {"mean", "std"}
    w = w.half().eval()
    x = gs["cold9_64x96/x"].half()
    with torch.no_grad():
        lat = w.encode_to_latent(torch.cat([x, x.flip(2)]))                 # batch of two different clips
    assert lat.shape == (2, 3, 16, 8, 12) and lat.dtype == torch.float32
    assert rel_l2(lat[0], gs["cold9_64x96/mu"][0].permute(1, 0, 2, 3)) < TOL
    assert not torch.allclose(lat[0], lat[1])


def test_decoder_built_and_cast_under_inference_mode(g):
    """A module constructed / cast inside torch.inference_mode() has parameters that track no version counter
    (``._version`` raises); the prepared-weights cache must not depend on it."""
    from realtime_video_b200.vae import VAEDecoderWrapper
    with torch.inference_mode():
        m = VAEDecoderWrapper()
        m.load_state_dict(synthetic_vae_params(seed=0), strict=False)
        m = m.half().eval()
        with pytest.raises(RuntimeError):
            m.decoder.conv1.weight._version
        px, cache = m(g["s8x12/z0"].half(), *([None] * 55))
        px2, _ = m(g["s8x12/z1"].half(), *cache)
    assert rel_l2(px, g["s8x12/px0"]) < TOL and rel_l2(px2, g["s8x12/px1"]) < TOL


@pytest.mark.parametrize("split", [(1, 1, 1, 1, 1), (2, 3), (4, 1), (5,)])
def test_decode_is_independent_of_how_the_stream_is_chunked(g, split):
    """Five latent frames decoded in one call, one by one, or in uneven calls: the causal decoder must produce the same
    17 pixel frames (1 + 4 * 4) — chunk boundaries (MAX_CHUNK = 3, first-frame special case, feature-cache roll, the
    up3d time_conv cache incl. its one-frame `where` quirk) may not show."""
    z = torch.cat([g["s8x12/z0"], g["s8x12/z1"]], dim=1)[:, :5].half()          # [1, 5, 16, 8, 12]

    def decode(parts):
        m = decoder()
        cache, outs, i = [None] * 55, [], 0
        with torch.no_grad():
            for n in parts:
                px, cache = m(z[:, i:i + n], *cache)
                outs.append(px)
                i += n
        return torch.cat(outs, dim=1)

    ref = decode((3, 2))                       # the server's pattern for these goldens: block 0, then the next frames
    got = decode(split)
    assert got.shape == ref.shape == (1, 17, 3, 64, 96)
    assert rel_l2(got, ref) < 1e-3, rel_l2(got, ref)
    assert rel_l2(ref[:, :9], g["s8x12/px0"]) < TOL


def test_two_interleaved_streams_on_one_decoder_have_value_semantics(g):
    """The reference's feature caches are values: one module can serve several streams whose callers keep their own
    cache lists (the server shares its models between sessions, release_server.py:760).  Here the lists alias the
    engine's buffers — copy-on-conflict must make that invisible: streams A and B interleaved on ONE module produce what
    two separate modules produce, and a single stream never pays a copy."""
    za0, za1, za2 = g["s8x12/z0"].half(), g["s8x12/z1"].half(), g["s8x12/z2"].half()
    zb0, zb1 = (za1[:, :3] * 0.5).contiguous(), (za0 * -0.7).contiguous()

    def alone(zs):
        m, cache, outs = decoder(), [None] * 55, []
        with torch.no_grad():
            for z in zs:
                px, cache = m(z, *cache)
                outs.append(px)
        return outs

    want_a, want_b = alone([za0, za1, za2]), alone([zb0, zb1])
    m = decoder()
    with torch.no_grad():
        a0, ca = m(za0, *([None] * 55))
        views = [t.data_ptr() for t in ca if t is not None]
        b0, cb = m(zb0, *([None] * 55))                        # stream B starts: A's list must become a snapshot
        assert all(t.data_ptr() != p for t, p in zip([t for t in ca if t is not None], views))
        a1, ca = m(za1, *ca)                                   # back to A: B's list snapshotted, A restored
        b1, cb = m(zb1, *cb)
        a2, ca = m(za2, *ca)
    for got, want in zip((a0, a1, a2), want_a):
        assert torch.equal(got, want)
    for got, want in zip((b0, b1), want_b):
        assert torch.equal(got, want)
    # single stream: the list handed out IS the buffers (no copy), and passing it back costs nothing
    m2 = decoder()
    with torch.no_grad():
        _, c1 = m2(za0, *([None] * 55))
        p1 = [t.data_ptr() for t in c1 if t is not None]
        _, c2 = m2(za1, *c1)
    assert [t.data_ptr() for t in c2 if t is not None] == p1 and [t.data_ptr() for t in c1 if t is not None] == p1


def test_interleaved_encoder_streams_have_value_semantics(g):
    from realtime_video_b200.vae import VAEEncoderWrapper
    torch.manual_seed(0)

    def encoder():
        from oracle.vae_oracle import synthetic_vae_params
        torch.manual_seed(0)                      # conv1 (WanVAE_.conv1) keeps its random init: same in every instance
        m = VAEEncoderWrapper()
        sd = synthetic_vae_params(seed=0)
        m.encoder.load_state_dict({k[len("encoder."):]: v for k, v in sd.items() if k.startswith("encoder.")}, strict=False)
        return m.half().eval()

    fa, fb = torch.rand(1, 3, 9, 64, 96).half() * 2 - 1, torch.rand(1, 3, 9, 64, 96).half() * 2 - 1

    def alone(f):
        m = encoder()
        with torch.no_grad():
            mu0, c = m(f[:, :, :5], [None] * 55, stream=False)
            mu1, c = m(f[:, :, 5:], c, stream=True)
        return mu0, mu1

    wa, wb = alone(fa), alone(fb)
    m = encoder()
    with torch.no_grad():
        a0, ca = m(fa[:, :, :5], [None] * 55, stream=False)
        b0, cb = m(fb[:, :, :5], [None] * 55, stream=False)
        a1, ca = m(fa[:, :, 5:], ca, stream=True)
        b1, cb = m(fb[:, :, 5:], cb, stream=True)
    assert torch.equal(a0, wa[0]) and torch.equal(a1, wa[1]) and torch.equal(b0, wb[0]) and torch.equal(b1, wb[1])
