"""Launch ONE of the model's GEMM shapes a few times (for `ncu --set full -k regex:gemm_f16 -s 2 -c 1`):
   python scripts/profile_gemm.py <qkv|out|ff1|ff2> [M] [bn]      (f16x3 path, operands as fp16 planes)"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from omnitokenizer_b200 import _cabi, layout as L

dev = torch.device("cuda:0")
shape = sys.argv[1] if len(sys.argv) > 1 else "ff1"
M = int(sys.argv[2]) if len(sys.argv) > 2 else 40960
_cabi.load()
if len(sys.argv) > 3:
    _cabi.set_option("f16_bn", int(sys.argv[3]))
g = torch.Generator(device=dev).manual_seed(1)
rnd = lambda *s: torch.rand(*s, device=dev, generator=g) - 0.5
if shape == "ff1":
    inner, ku = 1365, 1408
    A = rnd(M, 512); ah, al = L.split_f16(A)
    wh, wl = L.split_f16(L.pad_rows(L.pack_geglu(rnd(2 * inner, 512) * 0.1, inner, ku), 256))
    U = torch.empty(2, M, ku, dtype=torch.int16, device=dev)
    fn = lambda: _cabi.linear_h(a_hi=ah, a_lo=al, lda=512, w_hi=wh, w_lo=wl, u_hi=U[0], u_lo=U[1], ldu=ku, M=M, N=2 * ku, K=512, epilogue=_cabi.EPI_GEGLU)
elif shape == "ff2":
    A = rnd(M, 1408); ah, al = L.split_f16(A)
    wh, wl = L.split_f16(L.pad_rows(rnd(512, 1408) * 0.1, 256))
    X = rnd(M, 512)
    fn = lambda: _cabi.linear_h(a_hi=ah, a_lo=al, lda=1408, w_hi=wh, w_lo=wl, c=X, ldc=512, M=M, N=512, K=1408, residual=X, ldr=512, epilogue=_cabi.EPI_NONE)
elif shape == "out":
    A = rnd(M, 512); ah, al = L.split_f16(A)
    wh, wl = L.split_f16(L.pad_rows(rnd(512, 512) * 0.1, 256))
    X = rnd(M, 512)
    fn = lambda: _cabi.linear_h(a_hi=ah, a_lo=al, lda=512, w_hi=wh, w_lo=wl, c=X, ldc=512, M=M, N=512, K=512, residual=X, ldr=512, epilogue=_cabi.EPI_NONE)
else:  # qkv with rope + l2norm + scale
    A = rnd(M, 512); ah, al = L.split_f16(A); a2h, a2l = L.split_f16(rnd(M, 512))
    wh, wl = L.split_f16(L.pad_rows(rnd(1536, 512) * 0.1, 256))
    cos, sin = [t.to(dev) for t in L.rope_tables(1024, 64)]
    qs, ks = rnd(64) + 1.0, rnd(64) + 1.0
    C = torch.empty(M, 1536, device=dev)
    fn = lambda: _cabi.linear_h(a_hi=ah, a_lo=al, a2_hi=a2h, a2_lo=a2l, n_split=512, lda=512, w_hi=wh, w_lo=wl, c=C, ldc=1536, M=M, N=1536, K=512,
                                epilogue=_cabi.EPI_QKV, q_scale=qs, k_scale=ks, rope_cos=cos, rope_sin=sin, qk_cols=1024, tokens=1024)
flush = torch.zeros(64 * 1024 * 1024, device=dev)
for _ in range(4):
    flush.add_(1.0)
    fn()
torch.cuda.synchronize()
print("done")
