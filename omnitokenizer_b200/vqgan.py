"""Drop-in `OmniTokenizer_VQGAN`: the reference's module API and checkpoint layout over the
omnitok_b200 CUDA kernels.

Mirrors /root/reference/OmniTokenizer/omnitokenizer.py:63-768 for the inference surface that
vqgan_eval.py, lm_transformer.py, transformer_eval.py and the DiT/Latte trainers touch:
``__init__(args)``, ``add_model_specific_args``, ``load_state_dict`` / ``state_dict`` (same key names
and shapes, including the dead keys), ``load_from_checkpoint``, ``encode``, ``decode``,
``forward(x, log_image=True)`` and the attributes those callers poke.  The parameter containers
below define NO torch forward math -- every per-token operation runs in libomnitok_b200.so
(omnitokenizer_b200/engine.py); there is no CPU or eager fallback.

Out of scope: GAN training (discriminators, LPIPS, optimizers).  Their checkpoint
keys are reported as `unexpected_keys` by ``load_state_dict(strict=False)``, which is how
vqgan_eval.py:62-71 already loads.
"""
from __future__ import annotations

import argparse
import random
import math
from typing import Optional

import torch
import torch.nn as nn

from .engine import Engine


# --------------------------------------------------------------------------------------------
# parameter containers (names == reference state_dict keys)
# --------------------------------------------------------------------------------------------

class _Slot(nn.Module):
    """Parameter-free placeholder for the reference's Rearrange / GEGLU / Dropout entries so that
    nn.Sequential indices (and therefore state_dict keys) line up."""


class _GammaBetaNorm(nn.Module):          # modules/attention.py:73-80
    def __init__(self, dim):
        super().__init__()
        self.gamma = nn.Parameter(torch.ones(dim))
        self.register_buffer("beta", torch.zeros(dim))


class _PEG(nn.Module):                    # modules/attention.py:298-302
    def __init__(self, dim):
        super().__init__()
        self.dsconv = nn.Conv3d(dim, dim, 3, groups=dim)


class _ContinuousPositionBias(nn.Module):  # modules/attention.py:535-560 (dead under SDPA; keys kept)
    def __init__(self, dim, heads, layers=2):
        super().__init__()
        self.net = nn.ModuleList([])
        self.net.append(nn.Sequential(nn.Linear(2, dim), nn.LeakyReLU(0.1)))
        for _ in range(layers - 1):
            self.net.append(nn.Sequential(nn.Linear(dim, dim), nn.LeakyReLU(0.1)))
        self.net.append(nn.Linear(dim, heads))


class _Attention(nn.Module):              # modules/attention.py:342-393
    def __init__(self, dim, dim_head, heads, spatial_pos):
        super().__init__()
        inner = dim_head * heads
        if spatial_pos == "rel":
            self.spatial_rel_pos_bias = _ContinuousPositionBias(dim=dim, heads=heads)
        self.norm = _GammaBetaNorm(dim)
        self.context_norm = _GammaBetaNorm(dim)
        self.to_q = nn.Linear(dim, inner, bias=False)
        self.to_kv = nn.Linear(dim, inner * 2, bias=False)
        self.q_scale = nn.Parameter(torch.ones(dim_head))
        self.k_scale = nn.Parameter(torch.ones(dim_head))
        self.to_out = nn.Linear(inner, dim, bias=False)


class _WindowAttention(nn.Module):        # modules/attention.py:216-252
    def __init__(self, dim, window_size, heads):
        super().__init__()
        ws = window_size
        self.norm = _GammaBetaNorm(dim)
        self.relative_position_bias_table = nn.Parameter(torch.zeros((2 * ws - 1) * (2 * ws - 1), heads))
        coords = torch.stack(torch.meshgrid(torch.arange(ws), torch.arange(ws), indexing="ij")).flatten(1)
        rel = (coords[:, :, None] - coords[:, None, :]).permute(1, 2, 0).contiguous()
        rel[:, :, 0] += ws - 1
        rel[:, :, 1] += ws - 1
        rel[:, :, 0] *= 2 * ws - 1
        self.register_buffer("relative_position_index", rel.sum(-1))
        self.qkv = nn.Linear(dim, dim * 3, bias=False)
        self.proj = nn.Linear(dim, dim)
        nn.init.trunc_normal_(self.relative_position_bias_table, std=0.02)


def _feed_forward(dim, mult):             # modules/attention.py:159-168
    inner = int(mult * (2 / 3) * dim)
    return nn.Sequential(nn.LayerNorm(dim), nn.Linear(dim, inner * 2, bias=False), _Slot(), _Slot(),
                         nn.Linear(inner, dim, bias=False))


class _Transformer(nn.Module):            # modules/attention.py:588-652
    def __init__(self, dim, block, dim_head, heads, ff_mult, window_size, spatial_pos):
        super().__init__()
        self.layers = nn.ModuleList([])
        for blk in block:
            if blk == "t":
                self.layers.append(nn.ModuleList([_PEG(dim), _Attention(dim, dim_head, heads, spatial_pos), None,
                                                  _feed_forward(dim, ff_mult)]))
            elif blk == "w":
                self.layers.append(nn.ModuleList([None, _WindowAttention(dim, window_size, heads), None,
                                                  _feed_forward(dim, ff_mult)]))
            else:
                raise NotImplementedError(
                    f"transformer block {blk!r}: pooling ('a','m','l') / upsampling ('n','r') blocks are used by no "
                    f"shipped config and are not implemented")
        self.block = block
        self.norm_out = _GammaBetaNorm(dim)


def _pair(v):
    return v if isinstance(v, tuple) else (v, v)


def _check_patch_embed(a):
    if a.patch_embed not in ("linear", "cnn"):
        raise NotImplementedError(f"patch_embed={a.patch_embed!r}: the reference itself raises for anything but linear / cnn")
    if a.patch_embed == "cnn" and getattr(a, "norm_type", "batch") != "batch":
        # Normalize(image_channels=3, 'group') is GroupNorm(32 groups, 3 channels): the reference cannot even construct it
        raise NotImplementedError("patch_embed='cnn' needs --norm_type batch (GroupNorm(32, 3) is invalid in the reference too)")


class OmniTokenizer_Encoder(nn.Module):   # omnitokenizer.py:772-868
    def __init__(self, a):
        super().__init__()
        _check_patch_embed(a)
        self.image_size = _pair(a.resolution)
        self.patch_size = _pair(a.patch_size)
        self.temporal_patch_size = a.temporal_patch_size
        self.block = a.enc_block
        k = a.image_channels * a.patch_size * a.patch_size
        dim = a.embedding_dim
        p, pt = a.patch_size, a.temporal_patch_size
        if a.patch_embed == "cnn":      # omnitokenizer.py:823-838: Conv3d + Normalize('batch' = SyncBatchNorm) + Rearrange
            self.to_patch_emb_first_frame = nn.Sequential(nn.Conv3d(a.image_channels, dim, (1, p, p), stride=(1, p, p)),
                                                          nn.BatchNorm3d(dim), _Slot())
            self.to_patch_emb = nn.Sequential(nn.Conv3d(a.image_channels, dim, (pt, p, p), stride=(pt, p, p)),
                                              nn.BatchNorm3d(dim), _Slot())
        else:
            self.to_patch_emb_first_frame = nn.Sequential(_Slot(), nn.LayerNorm(k), nn.Linear(k, dim), nn.LayerNorm(dim))
            self.to_patch_emb = nn.Sequential(_Slot(), nn.LayerNorm(k * pt), nn.Linear(k * pt, dim), nn.LayerNorm(dim))
        kw = dict(dim=dim, dim_head=a.dim_head, heads=a.heads, ff_mult=a.ff_mult, window_size=a.twod_window_size)
        self.enc_spatial_transformer = _Transformer(block=a.enc_block, spatial_pos=a.spatial_pos, **kw)
        # temporal transformers are built without spatial_pos -> default "rel" -> dead bias-MLP keys
        self.enc_temporal_transformer = _Transformer(block="t" * a.temporal_depth, spatial_pos="rel", **kw)


class OmniTokenizer_Decoder(nn.Module):   # omnitokenizer.py:950-1035
    def __init__(self, a):
        super().__init__()
        self.image_size = _pair(a.resolution)
        self.patch_size = _pair(a.patch_size)
        self.block = a.dec_block
        k = a.image_channels * a.patch_size * a.patch_size
        dim = a.embedding_dim
        kw = dict(dim=dim, dim_head=a.dim_head, heads=a.heads, ff_mult=a.ff_mult, window_size=a.twod_window_size)
        self.dec_spatial_transformer = _Transformer(block=a.dec_block, spatial_pos=a.spatial_pos, **kw)
        self.dec_temporal_transformer = _Transformer(block="t" * a.temporal_depth, spatial_pos="rel", **kw)
        _check_patch_embed(a)
        p, pt = a.patch_size, a.temporal_patch_size
        if a.patch_embed == "cnn":      # omnitokenizer.py:1019-1035: Rearrange + ConvTranspose3d + Normalize(channels)
            self.to_pixels_first_frame = nn.Sequential(_Slot(), nn.ConvTranspose3d(dim, a.image_channels, (1, p, p),
                                                                                  stride=(1, p, p)), nn.BatchNorm3d(a.image_channels))
            self.to_pixels = nn.Sequential(_Slot(), nn.ConvTranspose3d(dim, a.image_channels, (pt, p, p), stride=(pt, p, p)),
                                           nn.BatchNorm3d(a.image_channels))
        else:
            self.to_pixels_first_frame = nn.Sequential(nn.Linear(dim, k), _Slot())
            self.to_pixels = nn.Sequential(nn.Linear(dim, k * pt), _Slot())


class Codebook(nn.Module):                # modules/codebook.py:11-28
    def __init__(self, n_codes, embedding_dim, no_random_restart=False, restart_thres=1.0, usage_sigma=0.99):
        super().__init__()
        self.register_buffer("embeddings", torch.randn(n_codes, embedding_dim))
        self.register_buffer("N", torch.zeros(n_codes))
        self.register_buffer("z_avg", self.embeddings.data.clone())
        self.register_buffer("codebook_usage", torch.zeros(n_codes))
        self.call_cnt = 0
        self.usage_sigma = usage_sigma
        self.n_codes = n_codes
        self.embedding_dim = embedding_dim
        self._need_init = True
        self.no_random_restart = no_random_restart
        self.restart_thres = restart_thres

    def dictionary_lookup(self, encodings):
        return torch.nn.functional.embedding(encodings, self.embeddings)


_BACKFILL = dict(twod_window_size=4, defer_temporal_pool=False, defer_spatial_pool=False, spatial_pos="rel",
                 logitslaplace_weight=0.0, gen_upscale=None, initialize_vit=False, use_vae=False, kl_weight=0.000001,
                 apply_diffaug=False, apply_noise=False, apply_blur=False, sigmoid_in_disc=False,
                 activation_in_disc="leaky_relu", video_perceptual_weight=0.0, grad_clip_val_disc=1.0,
                 disloss_check_thres=None, perloss_check_thres=None, recloss_check_thres=None, resolution_scale=None)


class OmniTokenizer_VQGAN(nn.Module):
    """B200-native stand-in for OmniTokenizer.OmniTokenizer_VQGAN (omnitokenizer.py:63)."""

    def __init__(self, args):
        super().__init__()
        self.args = args
        # same hasattr back-fills as the reference so old checkpoints' Namespaces construct (omnitokenizer.py:70-236)
        if not hasattr(args, "enc_block"):
            args.enc_block = "t" * args.spatial_depth
        if not hasattr(args, "dec_block"):
            args.dec_block = "t" * args.spatial_depth
        for k, v in _BACKFILL.items():
            if not hasattr(args, k):
                setattr(args, k, v)
        for k, v in dict(patch_embed="linear", attn_dropout=0.0, ff_dropout=0.0, ff_mult=4.0, dim_head=64, heads=8,
                         image_channels=3, use_external_codebook=False, l2_code=False, no_random_restart=False,
                         restart_thres=1.0, causal_in_temporal_transformer=False, causal_in_peg=False,
                         temporal_depth=4, sample_every_n_frames=1, downsample=(4, 4, 4)).items():
            if not hasattr(args, k):
                setattr(args, k, v)
        self.embedding_dim = args.embedding_dim
        self.n_codes = args.n_codes
        self.logitslaplace_weight = args.logitslaplace_weight
        self.gen_upscale = args.gen_upscale
        self.resolution = args.resolution
        self.patch_size = args.patch_size
        self.resolution_scale = args.resolution_scale
        if args.defer_temporal_pool or args.defer_spatial_pool or args.gen_upscale is not None:
            raise NotImplementedError("defer_*_pool / gen_upscale are multi-resolution training options that change the "
                                      "architecture (pooling / upsampling blocks), outside the encode/decode hot path "
                                      "(SURVEY.md 8f.4)")
        if args.use_external_codebook:
            raise NotImplementedError("--use_external_codebook (vendored lucidrains quantizers) is never set by the shipped "
                                      "scripts and is out of scope")
        self.encoder = OmniTokenizer_Encoder(args)
        self.decoder = OmniTokenizer_Decoder(args)
        self.use_vae = args.use_vae
        self.kl_weight = args.kl_weight
        self.codebook = Codebook(args.n_codes, args.codebook_dim, no_random_restart=args.no_random_restart,
                                 restart_thres=args.restart_thres)
        out = args.codebook_dim * 2 if self.use_vae else args.codebook_dim
        self.pre_vq_conv = nn.Sequential(_Slot(), nn.Linear(args.embedding_dim, out), _Slot())
        self.post_vq_conv = nn.Sequential(_Slot(), nn.Linear(args.codebook_dim, args.embedding_dim), _Slot())
        self.use_external_codebook = args.use_external_codebook
        self.l2_code = args.l2_code
        self.hparams = argparse.Namespace(args=args)
        self._engine: Optional[Engine] = None
        self._engine_key = None
        self.requires_grad_(False)      # inference module: the CUDA path has no autograd

    # ---------------------------------------------------------------- plumbing
    @property
    def device(self):
        return self.codebook.embeddings.device

    @property
    def latent_shape(self):               # omnitokenizer.py:239-245
        a = self.args
        inp = (a.sequence_length // a.sample_every_n_frames, a.resolution, a.resolution)
        return tuple(s // d for s, d in zip(inp, a.downsample))

    def _apply(self, fn, *a, **k):
        self._engine = None
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        self._engine = None
        return super().load_state_dict(state_dict, strict=strict, **kw)

    @classmethod
    def load_from_checkpoint(cls, path, strict: bool = False, map_location="cpu", **kw):
        """Lightning-style loader (download.py:49, README.md:66): ckpt['hyper_parameters']['args'] + ['state_dict']."""
        ckpt = torch.load(path, map_location=map_location, weights_only=False)
        model = cls(ckpt["hyper_parameters"]["args"])
        model.load_state_dict(ckpt["state_dict"], strict=strict)
        return model

    def engine(self) -> Engine:
        """Packed-weight engine; rebuilt when weights / device / VAE mode change."""
        dev = self.device
        if dev.type != "cuda":
            raise RuntimeError("OmniTokenizer_VQGAN (omnitok_b200) runs on a CUDA device only; move the module with "
                               ".cuda() first -- there is no CPU fallback")
        key = (dev, bool(self.use_vae), sum(p._version for p in self.parameters()),
               self.codebook.embeddings.data_ptr(), self.codebook.embeddings._version)
        if self._engine is None or self._engine_key != key:
            with torch.cuda.device(dev):
                self._engine = Engine(self, dev)
            self._engine_key = key
        return self._engine

    def prepare(self):
        """Pack weights now (otherwise done lazily on first encode/decode)."""
        self.engine()
        return self

    def _track_usage(self, counts, M):
        """Eval-time side effects of Codebook.forward (modules/codebook.py:122-140): batch usage from the fixed-size
        histogram the search kernel filled (replaces torch.unique, no host sync), EMA of codebook_usage, call_cnt."""
        cb = self.codebook
        usage = counts[:cb.n_codes].float() / M
        if cb.call_cnt == 0:
            cb.codebook_usage.data = usage
        else:
            cb.codebook_usage.data = cb.usage_sigma * cb.codebook_usage.data + (1 - cb.usage_sigma) * usage
        cb.call_cnt += 1
        return usage

    def _empty_encode(self, x, is_image, include_embeddings):
        """B == 0 (a surplus rank of a batch-sharded run): right-shaped empty results, no kernel launch."""
        a = self.args
        T = 1 if is_image else x.shape[2]
        Tp, h, w = 1 + (T - 1) // a.temporal_patch_size, x.shape[-2] // a.patch_size, x.shape[-1] // a.patch_size
        if self.use_vae:
            shape = (0, a.codebook_dim, h, w) if is_image else (0, a.codebook_dim, Tp, h, w)
            return torch.empty(shape, device=self.device)
        enc = torch.empty((0, Tp, h, w), dtype=torch.int64, device=self.device)
        if include_embeddings:
            return torch.empty((0, a.codebook_dim, Tp, h, w), device=self.device), enc
        return enc

    # ---------------------------------------------------------------- the hot path
    @torch.no_grad()
    def encode(self, x, is_image, include_embeddings=False):
        """omnitokenizer.py:247-266."""
        eng = self.engine()
        if x.shape[0] == 0:
            return self._empty_encode(x, is_image, include_embeddings)
        with torch.cuda.device(self.device):
            xv = x.unsqueeze(2) if is_image else x
            ws, (B, Tp, h, w) = eng.encode(xv.float(), "raw" if self.use_vae else "vq")
            if not self.use_vae:
                enc = ws.idx.view(B, Tp, h, w).clone()
                self._track_usage(ws.counts, ws.M)             # Codebook.forward runs inside encode() too
                if include_embeddings:
                    z = eng.z_view(ws)
                    e = eng.E[ws.idx]
                    st = (e - z) + z
                    return st.view(B, Tp, h, w, -1).permute(0, 4, 1, 2, 3).contiguous(), enc
                return enc
            hpar = eng.z_view(ws)                                                 # (M, 2*cd) moments
            c = hpar.shape[1] // 2
            hpar = hpar.view(B, Tp, h, w, 2 * c).permute(0, 4, 1, 2, 3)
            mean, logvar = hpar[:, :c], torch.clamp(hpar[:, c:], -30.0, 20.0)     # vae.py:7-8
            noise = torch.randn(mean.shape).to(device=self.device)                # CPU generator, vae.py:16
            z = mean + torch.exp(0.5 * logvar) * noise
            return z.squeeze(2) if is_image else z.contiguous()

    def _check_cnn_grid(self, h, w):
        # the cnn decoder's Rearrange pins h to image_size // patch_size (omnitokenizer.py:1021): other grids raise there
        if self.args.patch_embed == "cnn" and h != self.resolution // self.patch_size:
            raise ValueError(f"patch_embed='cnn' decodes only the configured resolution ({self.resolution}): "
                             f"token grid {h}x{w} != {self.resolution // self.patch_size}")

    def _decode_inputs(self, encodings, is_image):
        """The index / flat-index / VAE 4-D 'b c h w' / 5-D 'b t h w c' conventions of omnitokenizer.py:268-317
        -> (dims (B,T',h,w), idx [M] | None, zc [M, cd] | None)."""
        if not self.use_vae:
            enc = encodings
            if enc.ndim == 2:
                B = enc.shape[0]
                if is_image:
                    h = w = int(math.sqrt(enc.shape[1])); Tp = 1
                else:
                    h = w = self.resolution // self.patch_size; Tp = enc.shape[1] // (h * w)
            elif enc.ndim == 3 and is_image:                                   # (B, h, w) is not a reference form
                raise ValueError("image indices must be (B, h*w) or (B, 1, h, w)")
            else:
                B, Tp, h, w = enc.shape
            self._check_cnn_grid(h, w)
            idx = enc.reshape(-1).to(device=self.device, dtype=torch.int64)
            if B > 0:
                # F.embedding device-asserts on out-of-range indices (omnitokenizer.py:270); same here, without a host sync
                torch._assert_async(((idx >= 0) & (idx < self.codebook.n_codes)).all(),
                                    "decode: code index out of range [0, n_codes)")
            return (B, Tp, h, w), idx, None
        z = encodings.to(device=self.device, dtype=torch.float32)
        if is_image:
            if z.ndim == 3:
                B = z.shape[0]; h = w = int(math.sqrt(z.shape[1])); Tp = 1
                zc = z.reshape(B * h * w, -1)
            else:
                B, c, h, w = z.shape; Tp = 1
                zc = z.permute(0, 2, 3, 1).reshape(B * h * w, c)
        else:
            if z.ndim == 3:
                B = z.shape[0]; h = w = self.resolution // self.patch_size; Tp = z.shape[1] // (h * w)
                zc = z.reshape(B * Tp * h * w, -1)
            else:
                B, Tp, h, w, c = z.shape
                zc = z.reshape(B * Tp * h * w, c)
        self._check_cnn_grid(h, w)
        return (B, Tp, h, w), None, zc

    def _decode(self, encodings, is_image, u8=None):
        eng = self.engine()
        with torch.cuda.device(self.device):
            dims, idx, zc = self._decode_inputs(encodings, is_image)
            B, Tp, h, w = dims
            if B == 0:
                T, H, W = 1 + (Tp - 1) * self.args.temporal_patch_size, h * self.patch_size, w * self.patch_size
                if u8 is not None:
                    return torch.empty((0, T, H, W, self.args.image_channels), dtype=torch.uint8, device=self.device)
                video = torch.empty((0, self.args.image_channels, T, H, W), device=self.device)
                return video.squeeze(2) if is_image else video
            return eng.decode(dims, idx=idx, zc=zc, u8=u8)

    @torch.no_grad()
    def decode(self, encodings, is_image):
        """omnitokenizer.py:268-317 (index / flat-index / VAE 4-D 'b c h w' / 5-D 'b t h w c' conventions)."""
        video = self._decode(encodings, is_image)
        return video.squeeze(2) if is_image else video

    @torch.no_grad()
    def decode_u8(self, encodings, is_image, affine=(1.0, 0.5, 0.0, 1.0, 255.0)):
        """decode() fused with the consumers' uint8 conversion: returns (B, T, H, W, C) uint8 (T = 1 for images) =
        trunc(clamp(x * mul + add, lo, hi) * post), bit-identical to the torch expression on decode()'s result.
        Default affine: vqgan_eval.py:139,147-148 `(clamp(x_recons + 0.5, 0, 1) * 255).byte()` in 'b t h w c' order;
        (255, 128, 0, 255, 1): DiT sample_ddp.py:163.  The device->host copy is 4x smaller than the fp32 video."""
        return self._decode(encodings, is_image, u8=affine)

    @torch.no_grad()
    def forward(self, x, optimizer_idx=None, log_image=False):
        """omnitokenizer.py:330-413, inference form (log_image=True).  The training branches
        (optimizer_idx 0/1: GAN / perceptual losses) are out of scope."""
        if optimizer_idx is not None or not log_image:
            raise NotImplementedError("only forward(x, log_image=True) (the vqgan_eval.py call) is implemented; the GAN "
                                      "training step is out of scope")
        eng = self.engine()
        is_image = x.ndim == 4
        if self.resolution_scale is not None:
            # multi-resolution option (omnitokenizer.py:334-355): ONE random.choices draw per call picks the scale, every
            # frame is resized bilinearly (align_corners=True) before the encoder; x is returned at that resolution
            scale = random.choices(self.resolution_scale)[0]
            side = int(x.shape[-2] * scale)
            flat = x if is_image else x.permute(0, 2, 1, 3, 4).reshape(-1, x.shape[1], x.shape[3], x.shape[4])
            flat = torch.nn.functional.interpolate(flat.float(), size=(side, side), mode="bilinear", align_corners=True)
            x = flat if is_image else flat.reshape(x.shape[0], x.shape[2], x.shape[1], side, side).permute(0, 2, 1, 3, 4).contiguous()
        with torch.cuda.device(self.device):
            xv = (x.unsqueeze(2) if is_image else x).float()
            ws, dims = eng.encode(xv, "raw" if self.use_vae else "vq")
            B, Tp, h, w = dims
            M = ws.M
            vq_output = None
            if not self.use_vae:
                z = eng.z_view(ws).clone()
                idx, counts = ws.idx.clone(), ws.counts.clone()
                x_recon = eng.decode(dims, idx=idx, straight_through=True)           # decoder sees (e - z) + z
                zq = eng.zq_view(ws).clone()
                cb = self.codebook
                n_codes = cb.n_codes
                usage = self._track_usage(counts, M)                                  # codebook.py:54-72, 133-138
                e = eng.E[idx]
                commitment = 0.25 * torch.mean((z - e) ** 2)                           # codebook.py:93
                perplexity = torch.exp(-torch.sum(usage * torch.log(usage + 1e-10)))   # codebook.py:122-123
                avg_usage = (cb.codebook_usage.data > (1 / n_codes)).sum() / n_codes
                vq_output = dict(embeddings=zq.view(B, Tp, h, w, -1).permute(0, 4, 1, 2, 3).contiguous(),
                                 encodings=idx.view(B, Tp, h, w), commitment_loss=commitment, perplexity=perplexity,
                                 avg_usage=avg_usage, batch_usage=usage)
            else:
                hpar = eng.z_view(ws)
                c = hpar.shape[1] // 2
                hp5 = hpar.view(B, Tp, h, w, 2 * c).permute(0, 4, 1, 2, 3)
                mean, logvar = hp5[:, :c], torch.clamp(hp5[:, c:], -30.0, 20.0)
                noise = torch.randn(mean.shape).to(device=self.device)               # drawn BEFORE randint (:368 vs :401)
                z = mean + torch.exp(0.5 * logvar) * noise
                zc = z.permute(0, 2, 3, 4, 1).reshape(M, c).contiguous()
                x_recon = eng.decode(dims, zc=zc)
            if is_image:
                x_recon = x_recon.squeeze(2)
                frames, frames_recon = x, x_recon
            else:
                T = x.shape[2]
                frame_idx = torch.randint(0, T, [B]).to(self.device)                  # omnitokenizer.py:401 (one CPU RNG draw)
                ar = torch.arange(B, device=self.device)
                frames, frames_recon = x[ar, :, frame_idx], x_recon[ar, :, frame_idx]
            return frames, frames_recon, x, x_recon, vq_output

    # ---------------------------------------------------------------- CLI surface
    @staticmethod
    def add_model_specific_args(parent_parser):
        """Same flag set as omnitokenizer.py:694-768 (stacks after base.VQGAN's and VideoData's parsers)."""
        parser = argparse.ArgumentParser(parents=[parent_parser], add_help=False)
        A = parser.add_argument
        for name, typ, default in (("--lr_min", float, 0.), ("--warmup_steps", int, 0), ("--warmup_lr_init", float, 0.),
                                   ("--grad_accumulates", int, 1), ("--grad_clip_val", float, 1.0),
                                   ("--grad_clip_val_disc", float, 1.0), ("--disloss_check_thres", float, None),
                                   ("--perloss_check_thres", float, None), ("--recloss_check_thres", float, None),
                                   ("--kl_weight", float, 0.), ("--video_perceptual_weight", float, 0.),
                                   ("--activation_in_disc", str, "leaky_relu"), ("--logitslaplace_weight", float, 0.),
                                   ("--dis_warmup_steps", int, 0), ("--dis_lr_multiplier", float, 1.),
                                   ("--patch_size", int, 16), ("--gen_upscale", int, None), ("--enc_block", str, "tttt"),
                                   ("--dec_block", str, "tttt"), ("--twod_window_size", int, 4),
                                   ("--temporal_patch_size", int, 2), ("--spatial_depth", int, 4),
                                   ("--temporal_depth", int, 4), ("--dim_head", int, 64), ("--heads", int, 8),
                                   ("--attn_dropout", float, 0.), ("--ff_dropout", float, 0.), ("--ff_mult", float, 4.),
                                   ("--codebook_type", str, "vq"), ("--codebook_dim", int, None),
                                   ("--commitment_weight", float, 0.25)):
            A(name, type=typ, default=default)
        for name in ("--force_alternation", "--use_vae", "--initialize_vit", "--sigmoid_in_disc", "--apply_blur",
                     "--apply_noise", "--apply_diffaug", "--dis_minlr_multiplier", "--defer_temporal_pool",
                     "--defer_spatial_pool", "--causal_in_temporal_transformer", "--causal_in_peg",
                     "--use_external_codebook", "--fp32_quant", "--l2_code"):
            A(name, action="store_true")
        A("--recon_loss_type", type=str, default="l1", choices=["l1", "l2"])
        A("--patch_embed", type=str, default="linear", choices=["linear", "cnn", "pixelshuffle"])
        A("--spatial_pos", type=str, default="rel", choices=["rel", "rope"])
        A("--resolution_scale", default=None, nargs="+", type=float)
        return parser

    @staticmethod
    def add_base_model_args(parent_parser):
        """The flags the reference takes from base.VQGAN.add_model_specific_args (base.py:245-269) -- vqgan_eval.py:44
        stacks that parser first; provided here so scripts can run without the legacy CNN tokenizer module."""
        parser = argparse.ArgumentParser(parents=[parent_parser], add_help=False)
        A = parser.add_argument
        A("--embedding_dim", type=int, default=256); A("--n_codes", type=int, default=2048)
        A("--n_hiddens", type=int, default=240); A("--lr", type=float, default=3e-4)
        A("--downsample", nargs="+", type=int, default=(4, 4, 4)); A("--disc_channels", type=int, default=64)
        A("--disc_layers", type=int, default=3); A("--discriminator_iter_start", type=int, default=50000)
        A("--disc_loss_type", type=str, default="hinge", choices=["hinge", "vanilla"])
        A("--apply_allframes", action="store_true"); A("--image_gan_weight", type=float, default=1.0)
        A("--video_gan_weight", type=float, default=1.0); A("--l1_weight", type=float, default=4.0)
        A("--gan_feat_weight", type=float, default=0.0); A("--perceptual_weight", type=float, default=0.0)
        A("--i3d_feat", action="store_true"); A("--restart_thres", type=float, default=1.0)
        A("--no_random_restart", action="store_true")
        A("--norm_type", type=str, default="group", choices=["batch", "group"])
        A("--padding_type", type=str, default="replicate", choices=["replicate", "constant", "reflect", "circular"])
        return parser


VQGAN = OmniTokenizer_VQGAN   # the reference's class name inside omnitokenizer.py


def canonical_args(extra=()):
    """argparse Namespace of the canonical config used by every shipped eval script
    (scripts/recons/eval_video.sh:1-9)."""
    p = argparse.ArgumentParser()
    p = OmniTokenizer_VQGAN.add_base_model_args(p)
    p = OmniTokenizer_VQGAN.add_model_specific_args(p)
    for f, d in (("--resolution", 256), ("--sequence_length", 17), ("--image_channels", 3),
                 ("--sample_every_n_frames", 1)):
        p.add_argument(f, type=int, default=d)
    argv = ("--patch_embed linear --patch_size 8 --temporal_patch_size 4 --spatial_depth 4 --temporal_depth 4 "
            "--embedding_dim 512 --disc_layers 3 --enc_block ttww --dec_block tttt --twod_window_size 8 "
            "--causal_in_temporal_transformer --causal_in_peg --dim_head 64 --heads 8 --apply_noise --apply_blur "
            "--spatial_pos rope --n_codes 8192 --codebook_dim 8 --l2_code --commitment_weight 1.0 "
            "--no_random_restart --resolution 256 --sequence_length 17 --norm_type batch").split()
    return p.parse_args(argv + list(extra))
