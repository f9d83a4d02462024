"""GPU model-level parity through the reference-facing module API (OmniTokenizer_VQGAN.encode /
decode / forward) against the oracle and the committed golden vectors.
Bars (BASELINE.json north_star): code indices bit-exact, pixels within 1e-3 abs."""
import os

import pytest
import torch

from oracle import omni_oracle as oo
from oracle import weights as W
from tests.util import build_model, check_sub, golden_setup, load_golden

pytestmark = pytest.mark.gpu
PIX_TOL = 1e-3


def _math_modes():
    return [m for m in os.environ.get("OMT_TEST_MATH", "fp32,3xtf32,f16x3").split(",") if m]


VARIANTS = {"base": dict(OMT_ATTN_F16="0", OMT_STATIC_U="0", OMT_PEG_KERNEL="3", OMT_ATTN_CTAS="1"),
            # the shipped defaults: fp16-plane attention core with two CTAs per SM, cp.async PEG, statically scaled GEGLU planes
            "default": dict(),
            # the alternate shapes of the default kernels: one attention CTA per SM, two-accumulator FF2
            "alt": dict(OMT_ATTN_F16="1", OMT_STATIC_U="0", OMT_PEG_KERNEL="4", OMT_ATTN_CTAS="1")}


@pytest.fixture(autouse=True, params=[v for v in os.environ.get("OMT_TEST_VARIANTS", "base,default,alt").split(",") if v])
def _kernel_variant(request, monkeypatch):
    """every model-level test runs with the conservative kernel set, the shipped defaults and the alternate shapes of the default kernels
    (the sets only differ in f16x3 math)"""
    for k in ("OMT_ATTN_F16", "OMT_STATIC_U", "OMT_PEG_KERNEL", "OMT_ATTN_CTAS"):
        monkeypatch.delenv(k, raising=False)
    for k, v in VARIANTS[request.param].items():
        monkeypatch.setenv(k, v)
    yield


@pytest.mark.parametrize("math", _math_modes())
@pytest.mark.parametrize("name", ["img64", "vid5x64", "vid9x128_b2", "img256_cfg1", "cnn_vid5x64"])
def test_vq_encode_decode_matches_golden(cuda, name, math):
    fx = load_golden(name)
    cfg, sd, x = golden_setup(fx)
    m = build_model(cfg, sd, cuda, math)
    is_image = x.ndim == 4
    emb, idx = m.encode(x.to(cuda), is_image, include_embeddings=True)
    assert idx.dtype == torch.int64 and tuple(idx.shape) == tuple(fx["idx"].shape)
    mism = int((idx.cpu() != fx["idx"].long()).sum())
    assert mism == 0, f"{mism}/{idx.numel()} code indices differ from the reference ({math})"
    check_sub(fx["emb"], emb, 1e-5, "embeddings")
    rec = m.decode(idx, is_image)
    err = check_sub(fx["rec"], rec, PIX_TOL, "reconstruction")
    print(f"{name} [{math}]: idx mismatches 0/{idx.numel()}, max |dpixel| {err:.2e}")
    if is_image:   # flat (B, h*w) index convention, omnitokenizer.py:271-275
        rec2 = m.decode(idx.reshape(idx.shape[0], -1), True)
        assert torch.equal(rec2, rec)


@pytest.mark.parametrize("math", _math_modes())
@pytest.mark.parametrize("name", ["vae_vid5x64", "vae_img64"])
def test_vae_matches_golden(cuda, name, math):
    fx = load_golden(name)
    cfg, sd, x = golden_setup(fx)
    m = build_model(cfg, sd, cuda, math)
    is_image = x.ndim == 4
    _orig = torch.randn
    try:       # the reference draws the noise from the global CPU RNG (vae.py:16); inject the recorded draw
        torch.randn = lambda *a, **k: fx["noise"].clone()
        z = m.encode(x.to(cuda), is_image)
    finally:
        torch.randn = _orig
    check_sub(fx["z"], z, 1e-4, "vae latent")
    rec = m.decode(z if is_image else z.permute(0, 2, 3, 4, 1), is_image)
    check_sub(fx["rec"], rec, PIX_TOL, "vae reconstruction")


@pytest.mark.parametrize("math", _math_modes())
def test_intermediate_activations_vs_oracle(cuda, math):
    """Layer-by-layer: the engine's canonical buffer after each transformer vs the oracle's taps."""
    cfg = oo.Config()
    sd = W.make_state_dict(cfg, 5)
    x = W.synthetic_input((1, 3, 5, 64, 64), 77)
    m = build_model(cfg, sd, cuda, math)
    eng = m.engine()
    taps = {}
    with torch.no_grad():
        oo.encoder(sd, cfg, x, taps)
    ws, dims = eng.encode(x.to(cuda), "vq")
    got = ws.X.cpu().view(taps["encoder_out"].shape)
    err = (got - taps["encoder_out"]).abs().max().item()
    assert err < 2e-4, f"encoder output differs from oracle by {err:.2e}"


@pytest.mark.parametrize("math", _math_modes())
def test_forward_log_image(cuda, math):
    cfg = oo.Config()
    sd = W.make_state_dict(cfg, 0)
    m = build_model(cfg, sd, cuda, math)
    # image: full statistics vs golden
    fx = load_golden("img64")
    _, _, x = golden_setup(fx)
    fr, frr, xx, xr, vq = m(x.to(cuda), log_image=True)
    assert torch.equal(vq["encodings"].cpu(), fx["idx"].long())
    check_sub(fx["fwd_rec"], xr, PIX_TOL, "forward recon")
    for k in ("commitment_loss", "perplexity", "avg_usage"):
        assert abs(float(vq[k]) - float(fx["fwd"][k])) <= 1e-4 * max(1.0, abs(float(fx["fwd"][k]))), k
    assert int((vq["batch_usage"] > 0).sum()) == fx["fwd"]["batch_usage_nnz"]
    assert m.codebook.call_cnt == 1 and torch.equal(m.codebook.codebook_usage, vq["batch_usage"])
    assert fr is not None and frr.shape == fr.shape
    # video: random-frame gather consumes one CPU RNG draw like the reference (omnitokenizer.py:401)
    xv = W.synthetic_input((2, 3, 5, 64, 64), 5)
    torch.manual_seed(123)
    want_idx = torch.randint(0, 5, [2])
    torch.manual_seed(123)
    fr, frr, xx, xr, vq = m(xv.to(cuda), log_image=True)
    with torch.no_grad():
        o = oo.forward_log_image(sd, cfg, xv, frame_idx=want_idx)
    assert torch.equal(vq["encodings"].cpu(), o[4]["encodings"])
    assert (xr.cpu() - o[3]).abs().max().item() < PIX_TOL
    assert torch.equal(fr.cpu(), o[0]) and (frr.cpu() - o[1]).abs().max().item() < PIX_TOL
    assert m.codebook.call_cnt == 2
    # encode() runs Codebook.forward too, so it moves the eval-time usage statistics (modules/codebook.py:133-138)
    before = m.codebook.codebook_usage.clone()
    idx = m.encode(xv.to(cuda), False)
    assert m.codebook.call_cnt == 3
    usage = torch.bincount(idx.reshape(-1).cpu(), minlength=8192).float() / idx.numel()
    want = 0.99 * before.cpu() + 0.01 * usage
    assert (m.codebook.codebook_usage.cpu() - want).abs().max().item() < 1e-7


def test_shape_errors_match_reference(cuda):
    cfg = oo.Config()
    m = build_model(cfg, W.make_state_dict(cfg, 0), cuda, "fp32")
    with pytest.raises(AssertionError, match="divisible by temporal patch size"):
        m.encode(torch.zeros(1, 3, 6, 64, 64, device=cuda), False)
    with pytest.raises(ValueError):
        m.encode(torch.zeros(1, 3, 96, 96, device=cuda), True)       # reference raises a reshape error here too
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        build_model(cfg, W.make_state_dict(cfg, 0), torch.device("cpu")).encode(torch.zeros(1, 3, 64, 64), True)


@pytest.mark.parametrize("math", _math_modes())
def test_determinism_and_batch_independence(cuda, math):
    """Samples are independent (the property batch-sharding relies on): a shard's codes equal the
    corresponding rows of the full batch."""
    cfg = oo.Config()
    m = build_model(cfg, W.make_state_dict(cfg, 2), cuda, math)
    x = W.synthetic_input((3, 3, 5, 64, 64), 9).to(cuda)
    full = m.encode(x, False)
    again = m.encode(x, False)
    assert torch.equal(full, again)
    part = m.encode(x[1:2], False)
    assert torch.equal(part, full[1:2])
    rf, rp = m.decode(full, False), m.decode(part, False)
    assert torch.equal(rf[1:2], rp)
