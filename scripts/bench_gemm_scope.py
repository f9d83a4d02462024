"""A/B of gemm_tc2's remote-arrive scope (omt_set_option("tc_arrive_cta", 0|1)) on the model's GEMM shapes,
CUDA events, L2 flushed between launches, median of 5."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from omnitokenizer_b200 import _cabi, layout as L

dev = torch.device("cuda:0")
shapes = {"ff1": (2752, 512, _cabi.EPI_GEGLU), "ff2": (512, 1376, 0), "proj": (512, 512, 0), "qkv": (1536, 512, 0)}
flush = torch.zeros(64 * 1024 * 1024, device=dev)
for M in (40960, 5120):
    for name, (N, K, epi) in shapes.items():
        A = torch.randn(M, K, device=dev)
        w = L.pad_rows((torch.rand(N, K, device=dev) - 0.5) * 0.1, 128)
        hi = L.tf32_round(w); lo = (w - hi).contiguous()
        ldc = N if epi == 0 else N // 2
        C = torch.empty(M, ldc, device=dev)
        line = f"M={M:6d} {name:5s}"
        for scope in (0, 1):
            _cabi.set_option("tc_arrive_cta", scope)
            ts = []
            for i in range(7):
                flush.add_(1.0)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                _cabi.call("omt_linear", A, K, 0, 0, 0, hi, lo, C, ldc, 0, 0, 0, M, N, K, None, None, 0, epi, _cabi.MATH_3XTF32)
                b.record(); torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            t = sorted(ts[2:])[2]
            line += f" | {'cta    ' if scope else 'cluster'} {t*1e3:7.1f} us {2.0*M*N*K/(t*1e-3)/1e12:6.1f} TF/s"
        print(line, flush=True)
