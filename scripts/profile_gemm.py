"""Launch the dominant GEMM (FF1 + GEGLU, M=40960 N=2752 K=512) a few times for `ncu --set full`,
and time variants with CUDA events (not under ncu)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from omnitokenizer_b200 import _cabi, layout as L

dev = torch.device("cuda:0")
mode = sys.argv[1] if len(sys.argv) > 1 else "time"
M = int(os.environ.get("GEMM_M", 40960))

def make(N, K, geglu=False):
    w = (torch.rand(N, K, device=dev) - 0.5) * 0.1
    w = L.pad_rows(w, 128)
    hi = L.tf32_round(w)
    return hi, (w - hi).contiguous(), w

def run(A, lda, whi, wlo, C, ldc, N, K, epi, math, res=None):
    _cabi.call("omt_linear", A, lda, 0, 0, 0, whi, wlo, C, ldc, 0, 0, 0, M, N, K, None, res, ldc if res is not None else 0, epi, math)

shapes = {"ff1": (2752, 512, _cabi.EPI_GEGLU), "ff2": (512, 1376, 0), "qproj": (512, 512, 0), "kv": (1024, 512, 0)}
flush = torch.zeros(64 * 1024 * 1024, device=dev)
if mode == "ncu":
    N, K, epi = shapes[os.environ.get("GEMM_SHAPE", "ff1")]
    A = torch.randn(M, K, device=dev); C = torch.empty(M, N, device=dev)
    whi, wlo, w = make(N, K)
    for _ in range(3):
        run(A, K, whi, wlo, C, N if epi == 0 else N // 2, N, K, epi, _cabi.MATH_3XTF32)
    torch.cuda.synchronize()
    sys.exit(0)

res = {}
for bn in (128, 256, 2):
    if bn == 2:
        _cabi.set_option("tc_kernel", 2)
    else:
        _cabi.set_option("tc_kernel", 1)
        _cabi.set_option("tc_block_n", bn)
    for name, (N, K, epi) in shapes.items():
        A = torch.randn(M, K, device=dev); C = torch.empty(M, N, device=dev)
        whi, wlo, w = make(N, K)
        for math, mname in ((_cabi.MATH_3XTF32, "3xtf32"), (_cabi.MATH_TF32, "tf32"), (_cabi.MATH_FP32, "fp32")):
            if math == _cabi.MATH_FP32 and bn != 128:
                continue
            if bn == 2 and math != _cabi.MATH_3XTF32:
                continue
            ts = []
            for i in range(7):
                flush.add_(1.0)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                run(A, K, whi if math != _cabi.MATH_FP32 else w, wlo, C, N if epi == 0 else N // 2, N, K, epi, math)
                b.record(); torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            t = sorted(ts[2:])[2]
            tf = 2.0 * M * N * K / (t * 1e-3) / 1e12
            res[f"{name}/{mname}/bn{bn}"] = (round(t * 1e3, 1), round(tf, 1))
            print(f"{name:6s} {mname:7s} bn={bn}: {t*1e3:8.1f} us  {tf:7.1f} TFLOP/s (algorithmic)", flush=True)
