"""The block driver stand-in (harness/server_loop.py = the text-to-video subset of release_server.py:344-736) end
to end on the CPU at a tiny geometry, kernels replaced by the fp32 stand-ins of tests/cpu_ops_emulation.py:
KV-cache recompute over the clean context frames, 4 denoise passes with re-noising, VAE decode, the sliding
context window and the first-frame re-encode through the VAE encoder: the loop's invariants.  Equality with the
unmodified release_server.GenerationSession and the reference-executed golden: tests/test_reference_callers_cpu.py."""
import pytest
import torch

from tests import cpu_ops_emulation as emu


@pytest.fixture(autouse=True)
def cpu_ops(monkeypatch):
    import realtime_video_b200.dit as dit
    import realtime_video_b200.wan_wrapper as ww
    import realtime_video_b200.vae as vae
    for mod in (dit, ww, vae):
        monkeypatch.setattr(mod, "ops", emu)


def make(keep_first_frame, decode=True, seed=7, blocks=4):
    from realtime_video_b200 import factory
    import harness
    from harness import GenerateParams, GenerationSession
    tr = factory.synthetic_transformer(size="tiny", device="cpu", dtype=torch.bfloat16, seed=0, dim=256, ffn_dim=512,
                                       num_heads=2, num_layers=2, text_dim=128)
    models = harness.build_models(tr, vae_decoder=factory.synthetic_vae_decoder(device="cpu"), device="cpu",
                                  vae_encoder=factory.synthetic_vae_encoder(device="cpu"))
    calls = {"enc": 0}
    enc = models.vae_encoder
    orig = enc.forward

    def counted(*a, **k):
        calls["enc"] += 1
        return orig(*a, **k)

    enc.forward = counted
    pe = factory.synthetic_prompt_embeds(device="cpu", text_dim=128, tokens=16)
    params = GenerateParams(width=96, height=64, seed=seed, kv_cache_num_frames=3, num_blocks=blocks,
                            keep_first_frame=keep_first_frame)
    return GenerationSession(params, models, prompt_embeds=pe, device="cpu", decode=decode), calls


def test_blocks_shapes_cache_indices_and_first_frame_reencode():
    sess, calls = make(keep_first_frame=False)
    fs = (64 // 16) * (96 // 16)                       # tokens per latent frame
    p = sess.models.pipeline
    assert p.frame_seq_length == fs and p.kv_cache1[0]["k"].shape[1] == (3 + 3) * fs
    outs = []
    for b in range(4):
        px = sess.generate_block()
        # block 0 decodes 9 frames and skips the first 3 (release_server.py:722-723)
        assert px.shape == (1, 6 if b == 0 else 12, 3, 64, 96) and px.dtype == torch.float32
        assert torch.isfinite(px).all() and float(px.abs().max()) <= 1.0
        # the block was denoised at cache positions [ctx, ctx + 3) frames, ctx = min(frames so far, kv)
        ctx = min(3 * b, 3)
        assert [p.kv_cache1[0]["global_end_index"], p.kv_cache1[0]["local_end_index"]] == [(ctx + 3) * fs] * 2
        assert sess.current_start_frame == 3 * (b + 1) and sess.block_idx == b + 1
        outs.append(px)
    assert sess.generate_block() is None                # num_blocks reached
    # blocks 0, 1: first latent frame kept; blocks 2, 3 (window slid): re-encoded from the oldest cached pixel frame
    assert calls["enc"] == 2
    assert len(sess.frame_context_cache) == 1 + (3 - 1) * 4
    assert torch.count_nonzero(sess.all_latents[:, :12]) > 0

    keep, calls_keep = make(keep_first_frame=True)
    outs_keep = [keep.generate_block() for _ in range(4)]
    assert calls_keep["enc"] == 0
    assert torch.equal(outs[0], outs_keep[0]) and torch.equal(outs[1], outs_keep[1])      # same until the window slides
    assert not torch.equal(outs[2], outs_keep[2])


def test_same_seed_same_stream_and_seed_matters():
    a, _ = make(True, decode=False, seed=3, blocks=3)
    b, _ = make(True, decode=False, seed=3, blocks=3)
    c, _ = make(True, decode=False, seed=4, blocks=3)
    for _ in range(3):
        xa, xb, xc = a.generate_block(), b.generate_block(), c.generate_block()
        assert xa.shape == (1, 3, 16, 8, 12) and torch.equal(xa, xb) and not torch.equal(xa, xc)
