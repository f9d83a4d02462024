"""GPU parity at BASELINE.json's full sizes (cfg-2 / cfg-3 / cfg-4 shapes) and size-independent properties."""
import os

import pytest
import torch

from oracle import omni_oracle as oo
from oracle import weights as W
from tests.util import build_model

pytestmark = pytest.mark.gpu


def _oracle(sd, cfg, x):
    oo.USE_LIBRARY_OPS = True          # torch's fused CPU ops (same arithmetic spec; tests/test_oracle.py pins both forms)
    try:
        torch.set_num_threads(min(16, os.cpu_count() or 1))
        with torch.no_grad():
            idx = oo.encode(sd, cfg, x)
            rec = oo.decode(sd, cfg, idx, x.ndim == 4)
    finally:
        oo.USE_LIBRARY_OPS = False
    return idx, rec


MATHS = [m for m in os.environ.get("OMT_TEST_MATH_FULL", "3xtf32,f16x3").split(",") if m]
_ORACLE_CACHE = {}


def _oracle_cached(shape, cfg, sd, x):
    if shape not in _ORACLE_CACHE:
        _ORACLE_CACHE[shape] = _oracle(sd, cfg, x)
    return _ORACLE_CACHE[shape]


@pytest.mark.parametrize("math", MATHS)
@pytest.mark.parametrize("shape", [(1, 3, 33, 512, 512),      # cfg-4: long-sequence stress (N=4096, T'=9), one sample
                                   (16, 3, 256, 256),         # cfg-2 shape (images), 16 of the 64: M = 16384 rows, > 1 GEMM wave
                                   (1, 3, 17, 256, 256)])     # cfg-3 sample
def test_full_size_parity(cuda, shape, math):
    cfg = oo.Config()
    sd = W.make_state_dict(cfg, 11)
    x = W.synthetic_input(shape, 321)
    m = build_model(cfg, sd, cuda, math)
    is_image = x.ndim == 4
    idx = m.encode(x.to(cuda), is_image)
    rec = m.decode(idx, is_image)
    idx_o, rec_o = _oracle_cached(shape, cfg, sd, x)
    mism = int((idx.cpu() != idx_o).sum())
    err = float((rec.cpu() - rec_o).abs().max())
    print(f"{shape}: idx mismatches {mism}/{idx.numel()}, max |dpixel| {err:.2e}")
    assert mism == 0
    assert err <= 1e-3


@pytest.mark.parametrize("math", MATHS)
def test_cfg5_vae_sample(cuda, math):
    """cfg-5 (VAE mode, --use_vae) at full size, one 17x256x256 sample: latents and pixels vs the oracle with the same noise."""
    cfg = oo.Config(use_vae=True)
    sd = W.make_state_dict(cfg, 13)
    x = W.synthetic_input((1, 3, 17, 256, 256), 322)
    m = build_model(cfg, sd, cuda, math)
    noise = torch.randn((1, 8, 5, 32, 32), generator=torch.Generator().manual_seed(7))
    _orig = torch.randn
    try:       # the reference draws the noise from the global CPU RNG (vae.py:16); inject one fixed draw on both sides
        torch.randn = lambda *a, **k: noise.clone()
        z = m.encode(x.to(cuda), False)
    finally:
        torch.randn = _orig
    rec = m.decode(z.permute(0, 2, 3, 4, 1), False)
    oo.USE_LIBRARY_OPS = True
    try:
        with torch.no_grad():
            z_o = oo.encode(sd, cfg, x, noise=noise)
            rec_o = oo.decode(sd, cfg, z_o.permute(0, 2, 3, 4, 1), False)
    finally:
        oo.USE_LIBRARY_OPS = False
    ez, er = float((z.cpu() - z_o).abs().max()), float((rec.cpu() - rec_o).abs().max())
    print(f"cfg-5 sample [{math}]: max |dz| {ez:.2e}, max |dpixel| {er:.2e}")
    assert ez <= 1e-4 and er <= 1e-3


@pytest.mark.parametrize("math", MATHS)
def test_cfg3_batch_properties(cuda, math):
    """Full cfg-3 batch (8 x 17x256x256): deterministic, and every sample's codes / pixels are independent of
    its batch neighbours -- the property batch-sharding over GPUs relies on."""
    cfg = oo.Config()
    m = build_model(cfg, W.make_state_dict(cfg, 12), cuda, math)
    x = W.synthetic_input((8, 3, 17, 256, 256), 654).to(cuda)
    full = m.encode(x, False)
    assert torch.equal(full, m.encode(x, False))
    for s, e in ((0, 1), (3, 5), (7, 8)):
        assert torch.equal(m.encode(x[s:e], False), full[s:e])
    rec = m.decode(full, False)
    assert torch.equal(m.decode(full[2:4], False), rec[2:4])
    assert torch.isfinite(rec).all() and tuple(rec.shape) == tuple(x.shape)
    # flat (B, T'hw) index convention at the configured resolution (omnitokenizer.py:281-286)
    assert torch.equal(m.decode(full.reshape(8, -1), False), rec)
