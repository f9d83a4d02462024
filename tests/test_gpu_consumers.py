"""GPU parity of the callers either side of encode / decode: the vqgan_eval.py loop, the LM token wire
format and the DiT / Latte latent formats -- omnitokenizer_b200.consumers through the module API vs the oracle's restatement
of the same reference lines, plus the fused uint8 un-patchify against the torch expression on the fp32 reconstruction."""
import argparse
import os

import pytest
import torch

import omnitokenizer_b200 as ob
from omnitokenizer_b200 import consumers as C
from oracle import omni_oracle as oo
from oracle import weights as W
from tests.util import build_model

pytestmark = pytest.mark.gpu
MATHS = [m for m in os.environ.get("OMT_TEST_MATH", "3xtf32,f16x3").split(",") if m]


def disabled_train(self, mode=True):           # vqgan_eval.py:19-22
    return self


@pytest.mark.parametrize("math", MATHS)
def test_vqgan_eval_loop(cuda, math, monkeypatch):
    """vqgan_eval.py:42-86 (stacked parsers -> model -> load_state_dict(strict=False) of a checkpoint that still carries
    discriminator keys -> attribute pokes -> disabled_train) and :114-167 (loop over a loader, forward(log_image=True),
    clamp/255/byte frames, usage sum)."""
    monkeypatch.setenv("OMT_MATH", math)
    parser = argparse.ArgumentParser()
    parser = ob.OmniTokenizer_VQGAN.add_base_model_args(parser)           # base.VQGAN.add_model_specific_args in the script
    parser = ob.OmniTokenizer_VQGAN.add_model_specific_args(parser)
    for f, d in (("--resolution", 64), ("--sequence_length", 5), ("--image_channels", 3), ("--sample_every_n_frames", 1)):
        parser.add_argument(f, type=int, default=d)                        # VideoData.add_data_specific_args supplies these
    parser.add_argument("--inference_type", type=str, default="video")
    args = parser.parse_args(("--patch_embed linear --patch_size 8 --temporal_patch_size 4 --spatial_depth 4 --temporal_depth 4 "
                              "--embedding_dim 512 --enc_block ttww --dec_block tttt --twod_window_size 8 "
                              "--causal_in_temporal_transformer --causal_in_peg --dim_head 64 --heads 8 --spatial_pos rope "
                              "--n_codes 8192 --codebook_dim 8 --l2_code --no_random_restart --norm_type batch").split())
    cfg = oo.Config(resolution=64)
    sd = W.make_state_dict(cfg, 21)
    ckpt = dict(sd)
    ckpt["video_discriminator.main.0.weight"] = torch.zeros(4, 3, 4, 4, 4)      # dropped by the script (:64-66)
    ckpt["image_discriminator.main.0.weight"] = torch.zeros(4, 3, 4, 4)        # reported as unexpected, ignored
    vqgan = ob.OmniTokenizer_VQGAN(args)
    state = {k: v for k, v in ckpt.items() if "video_discriminator" not in k}
    res = vqgan.load_state_dict(state, strict=False)
    assert not res.missing_keys and res.unexpected_keys == ["image_discriminator.main.0.weight"]
    vqgan = vqgan.to(cuda)
    vqgan.encoder.image_size = (args.resolution, args.resolution)
    vqgan.decoder.image_size = (args.resolution, args.resolution)
    num_codes = vqgan.codebook.n_codes
    vqgan.codebook._need_init = False
    vqgan.train = disabled_train.__get__(vqgan)
    vqgan.eval()
    loader = [{"video": W.synthetic_input((2, 3, 5, 64, 64), 100 + i)} for i in range(2)]
    total_usage = torch.zeros(num_codes, device=cuda)
    want_usage = torch.zeros(num_codes)
    state_o = None
    for batch in loader:
        x = batch["video"]
        torch.manual_seed(5)
        fidx = torch.randint(0, 5, [2])
        torch.manual_seed(5)                                                  # forward draws the random frame on the CPU RNG
        x_recons, frames, vq_output = C.eval_step(vqgan, x.to(cuda), total_usage)
        with torch.no_grad():
            state_o = state_o or {"call_cnt": 0, "codebook_usage": torch.zeros(num_codes)}
            o = oo.forward_log_image(sd, cfg, x, frame_idx=fidx, usage_state=state_o)
        assert torch.equal(vq_output["encodings"].cpu(), o[4]["encodings"])
        assert (x_recons.cpu() - o[3]).abs().max().item() <= 1e-3
        fake = torch.clamp(x_recons.detach().cpu() + 0.5, 0, 1)               # the script's own expression (:139,147-148)
        ref_frames = (fake * 255).permute(0, 2, 3, 4, 1).contiguous().byte()
        assert torch.equal(frames.cpu(), ref_frames)
        assert (frames.cpu().int() - oo.to_u8(o[3]).int()).abs().max().item() <= 1   # vs the oracle's bytes: one grey level
        want_usage += o[4]["batch_usage"]
        # fused path: same bytes as the torch expression applied to this model's own fp32 reconstruction
        codes = vqgan.encode(x.to(cuda), False)
        rec = vqgan.decode(codes, False)
        u8 = vqgan.decode_u8(codes, False)
        assert torch.equal(u8, C._to_u8(rec, C.EVAL_U8))
        assert torch.equal(C.reconstruct_u8(vqgan, x.to(cuda)), u8)
    assert (total_usage.cpu() - want_usage).abs().max().item() < 1e-6
    assert int((total_usage > 0).sum()) == int((want_usage > 0).sum())
    # image branch of the script (:185-196): 4-D input, encodings / batch_usage read from vq_output
    xi = W.synthetic_input((3, 3, 64, 64), 7)
    x_recons, frames, vq_output = C.eval_step(vqgan, xi.to(cuda))
    with torch.no_grad():
        o = oo.forward_log_image(sd, cfg, xi)
    assert torch.equal(vq_output["encodings"].cpu(), o[4]["encodings"]) and tuple(frames.shape) == (3, 1, 64, 64, 3)
    assert torch.equal(vqgan.decode_u8(vq_output["encodings"], True), C._to_u8(vqgan.decode(vq_output["encodings"], True).unsqueeze(2), C.EVAL_U8))


@pytest.mark.parametrize("math", MATHS)
def test_lm_token_wire_format(cuda, math):
    """lm_transformer.py:258-268 / :433-434 through the module API vs the oracle."""
    cfg = oo.Config(resolution=64)
    sd = W.make_state_dict(cfg, 22)
    m = build_model(cfg, sd, cuda, math)
    x = W.synthetic_input((2, 3, 9, 64, 64), 31)                            # T' = 3 latent frames
    for n in (0, 2):
        emb, tgt = C.encode_to_z(m, x.to(cuda), False, n)
        with torch.no_grad():
            emb_o, tgt_o = oo.encode_to_z(sd, cfg, x, False, n)
        assert tgt.dtype == torch.int64 and torch.equal(tgt.cpu(), tgt_o)
        assert tuple(emb.shape) == tuple(emb_o.shape) and (emb.cpu() - emb_o).abs().max().item() < 1e-5
    emb, tgt = C.encode_to_z(m, x.to(cuda), False, 0)
    offset = 1000                                                           # class-conditional vocabulary in front of the codes
    ix = (tgt + offset).unsqueeze(-1)
    ix[0, 0, 0] = 5                                                         # a conditioning token sampled by mistake clamps to code 0
    rec = C.decode_tokens(m, ix, False, cond_stage_vocab_size=offset)
    with torch.no_grad():
        rec_o = oo.decode_tokens(sd, cfg, ix.cpu(), False, offset)
    assert (rec.cpu() - rec_o).abs().max().item() <= 1e-3
    xi = W.synthetic_input((2, 3, 64, 64), 32)
    emb, tgt = C.encode_to_z(m, xi.to(cuda), True)
    with torch.no_grad():
        emb_o, tgt_o = oo.encode_to_z(sd, cfg, xi, True)
    assert torch.equal(tgt.cpu(), tgt_o) and tuple(emb.shape) == (2, 1, 8, 8, 8)
    assert (C.decode_tokens(m, tgt, True).cpu() - oo.decode_tokens(sd, cfg, tgt_o, True)).abs().max().item() <= 1e-3


@pytest.mark.parametrize("math", MATHS)
def test_dit_latte_latent_formats(cuda, math):
    """DiT (train.py:242, sample_ddp.py:162-163) and Latte (train.py:215-217, sample_ddp.py:201-206) through the module API."""
    cfg = oo.Config(use_vae=True, resolution=64)
    sd = W.make_state_dict(cfg, 23)
    m = build_model(cfg, sd, cuda, math)
    _orig = torch.randn
    try:
        # DiT: images
        xi = W.synthetic_input((2, 3, 64, 64), 41)
        noise = _orig((2, 8, 1, 8, 8), generator=torch.Generator().manual_seed(3))
        torch.randn = lambda *a, **k: noise.clone()                           # the posterior noise is a CPU-generator draw (vae.py:16)
        z = C.dit_encode_latents(m, xi.to(cuda))
        with torch.no_grad():
            z_o, img_o = oo.dit_roundtrip(sd, cfg, xi, noise)
        assert tuple(z.shape) == (2, 8, 8, 8) and (z.cpu() - z_o).abs().max().item() < 1e-4
        img = C.dit_decode_latents(m, z)
        assert img.dtype == torch.uint8 and tuple(img.shape) == (2, 64, 64, 3)
        assert (img.cpu().int() - img_o.int()).abs().max().item() <= 1
        fp = C.dit_decode_latents(m, z, as_uint8=False)
        assert torch.equal(img, C._to_u8(fp.unsqueeze(2), C.DIT_U8)[:, 0])   # fused bytes == the script's expression on fp32
        # Latte: clips 'b f c h w'
        xv = W.synthetic_input((1, 3, 5, 64, 64), 42).permute(0, 2, 1, 3, 4).contiguous()
        noise = _orig((1, 8, 2, 8, 8), generator=torch.Generator().manual_seed(4))
        torch.randn = lambda *a, **k: noise.clone()
        zl = C.latte_encode_latents(m, xv.to(cuda))
        with torch.no_grad():
            zl_o, vid_o, u8_o = oo.latte_roundtrip(sd, cfg, xv, noise)
        assert tuple(zl.shape) == (1, 2, 8, 8, 8) and (zl.cpu() - zl_o).abs().max().item() < 1e-4
        vid = C.latte_decode_latents(m, zl, as_uint8=False)
        assert tuple(vid.shape) == (1, 5, 3, 64, 64) and (vid.cpu() - vid_o).abs().max().item() <= 1e-3
        u8 = C.latte_decode_latents(m, zl)
        assert tuple(u8.shape) == (1, 5, 64, 64, 3) and torch.equal(u8, C._to_u8(vid.permute(0, 2, 1, 3, 4), C.EVAL_U8))
        assert (u8.cpu().int() - u8_o.int()).abs().max().item() <= 1
    finally:
        torch.randn = _orig


def test_unpatchify_u8_kernel(cuda):
    """omt_unpatchify_u8 alone: every byte equals the torch expression, both affines, first-frame and rest-frames rows."""
    from omnitokenizer_b200 import _cabi
    _cabi.load()
    B, T, H = 2, 5, 32
    g = torch.Generator().manual_seed(9)
    for first, rows, K in ((1, B * 16, 192), (0, B * 16, 768)):
        P = (torch.rand((rows, K), generator=g) - 0.5) * 3.0
        for aff in (C.EVAL_U8, C.DIT_U8):
            vid = torch.zeros((B, 3, T, H, H), device=cuda)
            out = torch.zeros((B, T, H, H, 3), dtype=torch.uint8, device=cuda)
            _cabi.call("omt_unpatchify", P.to(cuda), vid, B, 3, T, H, H, 8, 4, first)
            _cabi.call("omt_unpatchify_u8", P.to(cuda), out, B, 3, T, H, H, 8, 4, first, *aff)
            want = C._to_u8(vid, aff)
            sl = slice(0, 1) if first else slice(1, None)
            assert torch.equal(out[:, sl], want[:, sl])
