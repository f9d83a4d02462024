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


@pytest.mark.parametrize("shape", [(1, 3, 33, 512, 512),      # cfg-4: long-sequence stress (N=4096, T'=9), one sample
                                   (4, 3, 256, 256),          # cfg-2 shape (images), 4 of the 64
                                   (1, 3, 17, 256, 256)])     # cfg-3 sample
def test_full_size_parity(cuda, shape):
    cfg = oo.Config()
    sd = W.make_state_dict(cfg, 11)
    x = W.synthetic_input(shape, 321)
    m = build_model(cfg, sd, cuda, "3xtf32")
    is_image = x.ndim == 4
    idx = m.encode(x.to(cuda), is_image)
    rec = m.decode(idx, is_image)
    idx_o, rec_o = _oracle(sd, cfg, x)
    mism = int((idx.cpu() != idx_o).sum())
    err = float((rec.cpu() - rec_o).abs().max())
    print(f"{shape}: idx mismatches {mism}/{idx.numel()}, max |dpixel| {err:.2e}")
    assert mism == 0
    assert err <= 1e-3


def test_cfg3_batch_properties(cuda):
    """Full cfg-3 batch (8 x 17x256x256): deterministic, and every sample's codes / pixels are independent of
    its batch neighbours -- the property batch-sharding over GPUs relies on."""
    cfg = oo.Config()
    m = build_model(cfg, W.make_state_dict(cfg, 12), cuda, "3xtf32")
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
