"""Same-box A/B of the f16 attention core: time omt_attn_spatial_h of several builds of the library (paths on the command
line, default: the in-tree one) at the cfg-3 (40 sequences x 1024 tokens) and cfg-4 (9 x 4096) shapes, both kernel shapes
(attn_f16_ctas 1 | 2); CUDA events, L2 flushed between launches, median of 8; outputs are cross-checked between builds."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from omnitokenizer_b200 import _cabi
from omnitokenizer_b200 import layout as L

paths = sys.argv[1:] or [_cabi.lib_path()]
dev = torch.device("cuda:0")
flush = torch.zeros(64 * 1024 * 1024, device=dev)
libs = []
for p in paths:
    lib = ctypes.CDLL(os.path.abspath(p))
    res, args = _cabi.SIGNATURES["omt_attn_spatial_h"]
    lib.omt_attn_spatial_h.restype, lib.omt_attn_spatial_h.argtypes = res, args
    lib.omt_set_option.restype, lib.omt_set_option.argtypes = ctypes.c_int, [ctypes.c_char_p, ctypes.c_int]
    lib.omt_last_error.restype = ctypes.c_char_p
    libs.append((os.path.basename(p), lib))

for nseq, N in ((40, 1024), (9, 4096)):
    H, M = 8, nseq * N
    g = torch.Generator().manual_seed(N)
    q = torch.nn.functional.normalize(torch.randn(M, H, 64, generator=g), dim=-1)
    k = torch.nn.functional.normalize(torch.randn(M, H, 64, generator=g), dim=-1)
    v = torch.randn(M, H, 64, generator=g)
    qs = L.pow2_scale(1.0)
    def planes(t):
        x = t.reshape(M, 512) * qs
        hi = x.half()
        return hi.view(torch.int16).to(dev), (x - hi.float()).half().view(torch.int16).to(dev)
    qh, ql = planes(q); kh, kl = planes(k)
    vh, vl, vinv = L.split_rows_rs(v.reshape(M * H, 64))
    vh, vl = vh.reshape(M, 512).to(dev), vl.reshape(M, 512).to(dev)
    vinv = vinv.reshape(M, H).t().contiguous().to(dev)
    flops = 4.0 * N * N * 64 * H * nseq
    ref = None
    for name, lib in libs:
        for ctas in (1, 2):
            if lib.omt_set_option(b"attn_f16_ctas", ctas) != 0:
                if ctas == 2:
                    continue
            o = torch.full((M, 512), float("nan"), device=dev)
            st = torch.cuda.current_stream().cuda_stream
            def call():
                rc = lib.omt_attn_spatial_h(qh.data_ptr(), ql.data_ptr(), 512, kh.data_ptr(), kl.data_ptr(), 512, vh.data_ptr(),
                                            vl.data_ptr(), 512, vinv.data_ptr(), qs * qs, o.data_ptr(), None, None, 512, nseq, N, H,
                                            8.0, st)
                assert rc == 0, lib.omt_last_error()
            ts = []
            for i in range(10):
                flush.add_(1.0)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); call(); b.record(); torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            t = sorted(ts[2:])[4]
            if ref is None:
                ref = o.clone()
            err = float((o - ref).abs().max())
            print(f"N={N:5d} nseq={nseq:3d} {name:28s} ctas/SM={ctas}: {t*1e3:7.1f} us  {flops/t/1e9:6.1f} TFLOP/s algorithmic   "
                  f"max|o - first build| {err:.2e}", flush=True)
