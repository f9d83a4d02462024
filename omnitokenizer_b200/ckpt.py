"""Checkpoint tooling around the hot path.

`inflate_gen` turns an image-stage checkpoint (only the first-frame patch embedding / un-embedding trained) into the
initialisation of the joint image-video stage: the temporal-patch weights are synthesised from the first-frame ones
(OmniTokenizer/utils.py:11-75, called from vqgan_train.py:54).  It is plain state-dict surgery on the keys of the
reference layout, so a checkpoint inflated here loads into the reference and vice versa.
"""
from __future__ import annotations

from typing import Dict

import torch

_FIRST_E, _VIDEO_E = "encoder.to_patch_emb_first_frame", "encoder.to_patch_emb"
_FIRST_D, _VIDEO_D = "decoder.to_pixels_first_frame.0", "decoder.to_pixels.0"


def inflate_gen(state_dict: Dict[str, torch.Tensor], temporal_patch_size: int, spatial_patch_size: int = 8,
                strategy: str = "average", inflation_pe: bool = False) -> Dict[str, torch.Tensor]:
    """Image checkpoint -> video-capable checkpoint.  `strategy`: "average" spreads each first-frame weight evenly over the
    `temporal_patch_size` frames of a video patch (a static clip then embeds like its frame); "first" keeps the weights on
    the first frame of the patch and zeros the rest.  Layout: the video patch vector is (c, pt, p1, p2) flattened with c
    slowest, but the reference concatenates whole first-frame vectors `pt` times -- reproduced as is (utils.py:26-33)."""
    if strategy not in ("average", "first"):
        raise NotImplementedError(f"inflate_gen strategy {strategy!r} (utils.py:47-48)")
    pt = temporal_patch_size
    out = dict(state_dict)

    def tile(t: torch.Tensor, dim: int = 0) -> torch.Tensor:
        if strategy == "average":
            return torch.cat([t / pt] * pt, dim=dim)
        return torch.cat([t] + [torch.zeros_like(t)] * (pt - 1), dim=dim)

    sd = state_dict
    out[_VIDEO_E + ".1.weight"] = tile(sd[_FIRST_E + ".1.weight"])              # LayerNorm over the patch vector
    out[_VIDEO_E + ".1.bias"] = tile(sd[_FIRST_E + ".1.bias"])
    out[_VIDEO_E + ".2.weight"] = tile(sd[_FIRST_E + ".2.weight"], dim=-1)      # Linear(patch -> dim): columns
    out[_VIDEO_E + ".2.bias"] = sd[_FIRST_E + ".2.bias"]
    out[_VIDEO_E + ".3.weight"] = sd[_FIRST_E + ".3.weight"]                    # LayerNorm(dim)
    out[_VIDEO_E + ".3.bias"] = sd[_FIRST_E + ".3.bias"]
    out[_VIDEO_D + ".weight"] = tile(sd[_FIRST_D + ".weight"])                  # Linear(dim -> patch): rows
    out[_VIDEO_D + ".bias"] = tile(sd[_FIRST_D + ".bias"])
    return out
