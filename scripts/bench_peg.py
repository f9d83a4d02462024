"""Time the two tiled PEG kernels (omt_set_option("peg_kernel", 3|4)) at the cfg-3 shapes with CUDA events
(L2 flushed between launches) and check that they agree bit for bit."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from omnitokenizer_b200 import _cabi

dev = torch.device("cuda:0")
B = int(os.environ.get("PEG_B", 8))
T, h, w, C = 5, 32, 32, 512
M = B * T * h * w
x = torch.rand(M, C, device=dev) - 0.5
w27 = (torch.rand(27, C, device=dev) - 0.5) * 0.3
bias = (torch.rand(C, device=dev) - 0.5) * 0.1
flush = torch.zeros(64 * 1024 * 1024, device=dev)
out = {}
for temporal in (0, 1):
    for pk in (3, 4):
        _cabi.set_option("peg_kernel", pk)
        y = torch.empty_like(x)
        ts = []
        for i in range(8):
            flush.add_(1.0)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            _cabi.call("omt_peg_volume", x, y, w27, bias, B, T, h, w, C, temporal, 1)
            b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        t = sorted(ts[2:])[len(ts[2:]) // 2]
        out[(temporal, pk)] = y
        print(f"peg_kernel={pk} temporal={temporal} B={B}: {t*1e3:7.1f} us   {2 * M * C * 4 / t / 1e6:7.1f} GB/s (read+write once)", flush=True)
    print("bit-identical:", torch.equal(out[(temporal, 3)], out[(temporal, 4)]), flush=True)
