"""TEST INFRASTRUCTURE ONLY -- deterministic synthetic checkpoint in the reference's state_dict layout.

No checkpoints are available offline, so parity runs on seeded random weights.
Everything is drawn with ``torch.rand`` on a CPU generator and shaped with exact arithmetic
(adds / multiplies by constants), so the same (cfg, seed) gives the same bits on every host --
the golden vectors in tests/golden/ depend on that; ``fingerprint`` is stored with them.

Key names / shapes follow SURVEY.md Appendix B (reference files omnitokenizer.py:772-1118,
modules/attention.py, modules/codebook.py).  Scales, LN affines, biases and the window bias
table are deliberately away from their ones/zeros init so that a kernel ignoring them fails.
"""
from __future__ import annotations

import math
from typing import Dict

import torch

from .omni_oracle import Config


def make_state_dict(cfg: Config, seed: int = 0) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    def uni(shape, lo, hi):
        return torch.rand(shape, generator=g) * (hi - lo) + lo

    def linear(name, n_out, n_in, bias=True, gain=1.0):
        b = gain / math.sqrt(n_in)
        sd[name + ".weight"] = uni((n_out, n_in), -b, b)
        if bias:
            sd[name + ".bias"] = uni((n_out,), -b, b)

    def ln(name, dim, wname="weight", bname="bias", zero_beta=False):
        sd[f"{name}.{wname}"] = uni((dim,), 0.5, 1.5)
        sd[f"{name}.{bname}"] = torch.zeros(dim) if zero_beta else uni((dim,), -0.1, 0.1)

    C, H, D = cfg.embedding_dim, cfg.heads, cfg.dim_head
    inner = cfg.ff_inner
    p, pt, ch = cfg.patch_size, cfg.temporal_patch_size, cfg.image_channels
    k1, k2 = ch * p * p, ch * p * p * pt

    def bn(name, dim):
        sd[name + ".weight"] = uni((dim,), 0.5, 1.5)
        sd[name + ".bias"] = uni((dim,), -0.1, 0.1)
        sd[name + ".running_mean"] = uni((dim,), -0.2, 0.2)
        sd[name + ".running_var"] = uni((dim,), 0.5, 1.5)
        sd[name + ".num_batches_tracked"] = torch.tensor(1000, dtype=torch.int64)

    cnn = cfg.patch_embed == "cnn"
    for pre, k, PT in (("encoder.to_patch_emb_first_frame", k1, 1), ("encoder.to_patch_emb", k2, pt)):
        if cnn:       # omnitokenizer.py:823-838
            b = 1.0 / math.sqrt(k)
            sd[pre + ".0.weight"] = uni((C, ch, PT, p, p), -b, b)
            sd[pre + ".0.bias"] = uni((C,), -b, b)
            bn(pre + ".1", C)
        else:
            ln(pre + ".1", k)
            linear(pre + ".2", C, k)
            ln(pre + ".3", C)

    def t_layer(lp, temporal):
        b = 1.0 / math.sqrt(27)
        sd[lp + ".0.dsconv.weight"] = uni((C, 1, 3, 3, 3), -b, b)
        sd[lp + ".0.dsconv.bias"] = uni((C,), -b, b)
        a = lp + ".1"
        sd[a + ".q_scale"] = uni((D,), 0.5, 1.5)
        sd[a + ".k_scale"] = uni((D,), 0.5, 1.5)
        if temporal or cfg.spatial_pos == "rel":      # dead ContinuousPositionBias keys (Appendix B)
            linear(a + ".spatial_rel_pos_bias.net.0.0", C, 2)
            linear(a + ".spatial_rel_pos_bias.net.1.0", C, C)
            linear(a + ".spatial_rel_pos_bias.net.2", H, C)
        ln(a + ".norm", C, "gamma", "beta")           # beta is a buffer, zeros in real ckpts; honoured anyway
        ln(a + ".context_norm", C, "gamma", "beta")   # dead
        linear(a + ".to_q", H * D, C, bias=False)
        linear(a + ".to_kv", 2 * H * D, C, bias=False)
        linear(a + ".to_out", C, H * D, bias=False)
        ff(lp + ".3")

    def w_layer(lp):
        a = lp + ".1"
        ws = cfg.twod_window_size
        sd[a + ".relative_position_bias_table"] = uni(((2 * ws - 1) ** 2, H), -1.0, 1.0)
        coords = torch.stack(torch.meshgrid(torch.arange(ws), torch.arange(ws), indexing="ij")).flatten(1)
        rel = (coords[:, :, None] - coords[:, None, :]).permute(1, 2, 0).contiguous()
        rel[:, :, 0] += ws - 1
        rel[:, :, 1] += ws - 1
        rel[:, :, 0] *= 2 * ws - 1
        sd[a + ".relative_position_index"] = rel.sum(-1)
        ln(a + ".norm", C, "gamma", "beta")
        linear(a + ".qkv", 3 * C, C, bias=False)
        linear(a + ".proj", C, C)
        ff(lp + ".3")

    def ff(fp):
        ln(fp + ".0", C)
        linear(fp + ".1", inner * 2, C, bias=False)
        linear(fp + ".4", C, inner, bias=False)

    def tr(pre, block, temporal):
        for i, blk in enumerate(block):
            if blk == "t":
                t_layer(f"{pre}.layers.{i}", temporal)
            elif blk == "w":
                w_layer(f"{pre}.layers.{i}")
            else:
                raise NotImplementedError(blk)
        ln(pre + ".norm_out", C, "gamma", "beta")

    tr("encoder.enc_spatial_transformer", cfg.enc_block, False)
    tr("encoder.enc_temporal_transformer", "t" * cfg.temporal_depth, True)
    tr("decoder.dec_spatial_transformer", cfg.dec_block, False)
    tr("decoder.dec_temporal_transformer", "t" * cfg.temporal_depth, True)
    if cnn:           # omnitokenizer.py:1019-1035: ConvTranspose3d weight is (in=dim, out=channels, kt, kh, kw)
        for pre, PT in (("decoder.to_pixels_first_frame", 1), ("decoder.to_pixels", pt)):
            b = 1.0 / math.sqrt(C)
            sd[pre + ".1.weight"] = uni((C, ch, PT, p, p), -b, b)
            sd[pre + ".1.bias"] = uni((ch,), -b, b)
            bn(pre + ".2", ch)
    else:
        linear("decoder.to_pixels_first_frame.0", k1, C)
        linear("decoder.to_pixels.0", k2, C)
    # codebook ~ N(0,1)-ish (Irwin-Hall of 12 uniforms: exact adds, no transcendental)
    E = torch.rand((cfg.n_codes, cfg.codebook_dim, 12), generator=g).sum(-1) - 6.0
    sd["codebook.embeddings"] = E
    sd["codebook.N"] = torch.zeros(cfg.n_codes)
    sd["codebook.z_avg"] = E.clone()
    sd["codebook.codebook_usage"] = torch.zeros(cfg.n_codes)
    linear("pre_vq_conv.1", cfg.codebook_dim * (2 if cfg.use_vae else 1), C)
    linear("post_vq_conv.1", C, cfg.codebook_dim)
    return sd


def fingerprint(sd: Dict[str, torch.Tensor]) -> float:
    """Order-independent float64 checksum used to prove two hosts generated the same weights."""
    tot = 0.0
    for k in sorted(sd):
        v = sd[k].double()
        tot += float((v * torch.arange(1, v.numel() + 1, dtype=torch.float64).reshape(v.shape) % 7.0).sum())
    return tot


def synthetic_input(shape, seed: int = 1234) -> torch.Tensor:
    """uniform [-0.5, 0.5) fp32, the datasets' range (data.py:54)."""
    return torch.rand(shape, generator=torch.Generator().manual_seed(seed)) - 0.5
