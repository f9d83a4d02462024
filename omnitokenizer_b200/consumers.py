"""The callers on either side of encode / decode, as thin host functions over the module API:

* vqgan_eval.py's per-batch step (forward(x, log_image=True) -> clamp/255/uint8 frames + usage accounting),
* the autoregressive LM's token wire format (lm_transformer.py:258-268 encode_to_z, :433-434 decode of sampled tokens),
* the latent-diffusion consumers of the VAE variant (DiT / Latte: the 0.18215 latent scale and their layouts).

Each function names the reference lines it stands in for.  They work with any object exposing the reference's
OmniTokenizer_VQGAN API; with this package's module the uint8 conversions run fused in the un-patchify kernel
(omt_unpatchify_u8: the device->host copy shrinks 4x) instead of as torch elementwise passes over the fp32 video.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

LATENT_SCALE = 0.18215            # Diffusion/DiT/train.py:242, Diffusion/Latte/train.py:216

EVAL_U8 = (1.0, 0.5, 0.0, 1.0, 255.0)        # (clamp(x + 0.5, 0, 1) * 255).byte()     vqgan_eval.py:139,147-148; Latte sample_ddp.py:206
DIT_U8 = (255.0, 128.0, 0.0, 255.0, 1.0)     # clamp(255 * x + 128.0, 0, 255).to(uint8)  DiT sample_ddp.py:163


def _to_u8(video: torch.Tensor, affine) -> torch.Tensor:
    """torch form of the fused conversion: (B,C,T,H,W) fp32 -> (B,T,H,W,C) uint8, the reference's op order."""
    mul, add, lo, hi, post = affine
    t = torch.clamp(video * mul + add, lo, hi) * post
    return t.permute(0, 2, 3, 4, 1).contiguous().to(torch.uint8)


# ----------------------------------------------------------------------------------------------- vqgan_eval.py
@torch.no_grad()
def eval_step(vqgan, x: torch.Tensor, total_usage: Optional[torch.Tensor] = None):
    """One iteration of vqgan_eval.py's loops (:115-155 video, :185-196 image): forward(log_image=True), the
    reconstruction as the uint8 frames the FVD / FID feature extractors take ('b t h w c' = shift_dim(fake * 255, 1, -1)
    .byte(), :147-148), and the running codebook-usage sum (:150-152).  Returns (x_recons fp32, frames uint8, vq_output)."""
    _, _, _, x_recons, vq_output = vqgan(x, log_image=True)
    is_image = x.ndim == 4
    frames = _to_u8(x_recons.unsqueeze(2) if is_image else x_recons, EVAL_U8)
    if total_usage is not None and vq_output is not None:
        total_usage += vq_output["batch_usage"]
    return x_recons, frames, vq_output


@torch.no_grad()
def reconstruct_u8(vqgan, x: torch.Tensor) -> torch.Tensor:
    """encode -> decode with the eval script's uint8 conversion fused into the last kernel: (B,T,H,W,C) uint8 frames
    (T = 1 for images).  Equal to _to_u8(decode(encode(x)), EVAL_U8) byte for byte."""
    is_image = x.ndim == 4
    codes = vqgan.encode(x, is_image)
    if getattr(vqgan, "use_vae", False) and not is_image:
        codes = codes.permute(0, 2, 3, 4, 1)                # 'b c t h w' -> 'b t h w c' (omnitokenizer.py:313)
    if hasattr(vqgan, "decode_u8"):
        return vqgan.decode_u8(codes, is_image, EVAL_U8)
    rec = vqgan.decode(codes, is_image)
    return _to_u8(rec.unsqueeze(2) if is_image else rec, EVAL_U8)


# ----------------------------------------------------------------------------------------------- lm_transformer.py
@torch.no_grad()
def encode_to_z(vqgan, x: torch.Tensor, is_image: bool, sample_every_n_latent_frames: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
    """Net2NetTransformer.encode_to_z (lm_transformer.py:258-268): the GPT's view of a clip.
    Returns (embeddings channels-last (B, T'', h, w, C), targets int64 (B, T''*h*w)) where T'' keeps every n-th latent frame."""
    emb, targets = vqgan.encode(x, is_image, include_embeddings=True)
    if sample_every_n_latent_frames > 0:
        emb = emb[:, :, ::sample_every_n_latent_frames]
        targets = targets[:, ::sample_every_n_latent_frames]
    emb = emb.movedim(1, -1).contiguous()                     # shift_dim(x, 1, -1)
    return emb, targets.reshape(targets.shape[0], -1)


@torch.no_grad()
def decode_tokens(vqgan, ix: torch.Tensor, is_image: bool, cond_stage_vocab_size: int = 0,
                  first_stage_vocab_size: Optional[int] = None) -> torch.Tensor:
    """Sampled GPT tokens back to pixels (lm_transformer.py:433-434, :453-454): the class-conditional vocabulary offset is
    removed and the result clamped into the codebook, then the flat (B, T'hw) indices go through decode()."""
    if first_stage_vocab_size is None:
        first_stage_vocab_size = vqgan.codebook.n_codes
    index = torch.clamp(ix - cond_stage_vocab_size, min=0, max=first_stage_vocab_size - 1)
    if index.ndim == 3 and index.shape[-1] == 1:
        index = index.squeeze(-1)
    return vqgan.decode(index, is_image)


# ----------------------------------------------------------------------------------------------- DiT / Latte (VAE mode)
@torch.no_grad()
def dit_encode_latents(vae, x: torch.Tensor) -> torch.Tensor:
    """Diffusion/DiT/train.py:242: images (B,3,H,W) -> scaled latents (B,8,h,w)."""
    return vae.encode(x, is_image=True).mul_(LATENT_SCALE)


@torch.no_grad()
def dit_decode_latents(vae, samples: torch.Tensor, as_uint8: bool = True) -> torch.Tensor:
    """Diffusion/DiT/sample_ddp.py:162-163: latents (B,8,h,w) -> images; as_uint8: (B,H,W,3) uint8 =
    clamp(255 * x + 128.0, 0, 255), the array the script hands to PIL."""
    z = samples / LATENT_SCALE
    if not as_uint8:
        return vae.decode(z, is_image=True)
    if hasattr(vae, "decode_u8"):
        return vae.decode_u8(z, True, DIT_U8)[:, 0]
    return _to_u8(vae.decode(z, is_image=True).unsqueeze(2), DIT_U8)[:, 0]


@torch.no_grad()
def latte_encode_latents(vae, x_bfchw: torch.Tensor) -> torch.Tensor:
    """Diffusion/Latte/train.py:215-217: clips 'b f c h w' -> scaled latents 'b f c h w' (f = latent frames)."""
    x = x_bfchw.permute(0, 2, 1, 3, 4).contiguous()           # 'b f c h w -> b c f h w'
    z = vae.encode(x, is_image=False).mul_(LATENT_SCALE)
    return z.permute(0, 2, 1, 3, 4).contiguous()              # 'b c f h w -> b f c h w'


@torch.no_grad()
def latte_decode_latents(vae, samples_bfchw: torch.Tensor, as_uint8: bool = True) -> torch.Tensor:
    """Diffusion/Latte/sample/sample_ddp.py:201-206: latents 'b f c h w' -> 'b f h w c' -> decode(z / 0.18215).
    as_uint8: (B, F, H, W, 3) uint8 = (clamp(x + 0.5, 0, 1) * 255).byte(), the frames written to the .mp4;
    otherwise the fp32 video 'b f c h w' (:204)."""
    z = samples_bfchw.permute(0, 1, 3, 4, 2) / LATENT_SCALE   # 'b f c h w -> b f h w c', then the latent scale
    if as_uint8 and hasattr(vae, "decode_u8"):
        return vae.decode_u8(z, False, EVAL_U8)
    video = vae.decode(z, is_image=False)                     # 'b c f h w'
    if not as_uint8:
        return video.permute(0, 2, 1, 3, 4).contiguous()
    return _to_u8(video, EVAL_U8)
