"""CPU check of the host logic in omnitokenizer_b200.consumers (layouts, subsampling, vocabulary offsets, latent scale)
with an adapter that serves the module API from the oracle: every consumer function must reproduce the oracle's restatement
of the same reference lines exactly (same arithmetic underneath, so the comparison is bit for bit)."""
import torch

from omnitokenizer_b200 import consumers as C
from oracle import omni_oracle as oo
from oracle import weights as W


class _OracleBacked:
    """OmniTokenizer_VQGAN's call surface served by the CPU oracle (test adapter, no kernels)."""

    def __init__(self, cfg, sd, noise=None):
        self.cfg, self.sd, self.noise, self.use_vae = cfg, sd, noise, cfg.use_vae
        self.codebook = type("CB", (), {"n_codes": sd["codebook.embeddings"].shape[0]})()

    def encode(self, x, is_image, include_embeddings=False):
        with torch.no_grad():
            return oo.encode(self.sd, self.cfg, x, include_embeddings=include_embeddings, noise=self.noise)

    def decode(self, enc, is_image):
        with torch.no_grad():
            return oo.decode(self.sd, self.cfg, enc, is_image)

    def __call__(self, x, log_image=False):
        with torch.no_grad():
            return oo.forward_log_image(self.sd, self.cfg, x, frame_idx=torch.zeros(x.shape[0], dtype=torch.long))


def test_lm_and_eval_consumers_match_oracle():
    cfg = oo.Config(resolution=64)
    sd = W.make_state_dict(cfg, 22)
    m = _OracleBacked(cfg, sd)
    x = W.synthetic_input((1, 3, 9, 64, 64), 31)
    for n in (0, 2):
        emb, tgt = C.encode_to_z(m, x, False, n)
        emb_o, tgt_o = oo.encode_to_z(sd, cfg, x, False, n)
        assert torch.equal(emb, emb_o) and torch.equal(tgt, tgt_o)
    emb, tgt = C.encode_to_z(m, x, False)
    ix = (tgt + 77).unsqueeze(-1)
    assert torch.equal(C.decode_tokens(m, ix, False, 77), oo.decode_tokens(sd, cfg, ix, False, 77))
    total = torch.zeros(8192)
    rec, frames, vq = C.eval_step(m, x, total)
    assert torch.equal(frames, oo.to_u8(rec)) and torch.equal(total, vq["batch_usage"])
    assert torch.equal(C.reconstruct_u8(m, x), oo.to_u8(m.decode(m.encode(x, False), False)))


def test_vae_consumers_match_oracle():
    cfg = oo.Config(use_vae=True, resolution=64)
    sd = W.make_state_dict(cfg, 23)
    xi = W.synthetic_input((2, 3, 64, 64), 41)
    noise = torch.randn((2, 8, 1, 8, 8), generator=torch.Generator().manual_seed(3))
    m = _OracleBacked(cfg, sd, noise)
    z = C.dit_encode_latents(m, xi)
    z_o, img_o = oo.dit_roundtrip(sd, cfg, xi, noise)
    assert torch.equal(z, z_o) and torch.equal(C.dit_decode_latents(m, z), img_o)
    xv = W.synthetic_input((1, 3, 5, 64, 64), 42).permute(0, 2, 1, 3, 4).contiguous()
    noise = torch.randn((1, 8, 2, 8, 8), generator=torch.Generator().manual_seed(4))
    m = _OracleBacked(cfg, sd, noise)
    zl = C.latte_encode_latents(m, xv)
    zl_o, vid_o, u8_o = oo.latte_roundtrip(sd, cfg, xv, noise)
    assert torch.equal(zl, zl_o)
    assert torch.equal(C.latte_decode_latents(m, zl, as_uint8=False), vid_o)
    assert torch.equal(C.latte_decode_latents(m, zl), u8_o)
