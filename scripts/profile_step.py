"""One encode+decode step of the bench workload between cudaProfilerStart/Stop (for ncu --profile-from-start off)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

math = sys.argv[1] if len(sys.argv) > 1 else "3xtf32"
wl = sys.argv[2] if len(sys.argv) > 2 else "cfg3"
os.environ["OMT_MATH"] = math
dev = torch.device("cuda:0")
m = bench.make_model(dev)
shape = bench.WORKLOADS[wl]["shape"]
if os.environ.get("OMT_BENCH_BATCH"):
    shape = (int(os.environ["OMT_BENCH_BATCH"]),) + shape[1:]
x = (torch.rand(shape, generator=torch.Generator().manual_seed(1234)) - 0.5).to(dev)
is_image = len(shape) == 4
for _ in range(2):
    rec = m.decode(m.encode(x, is_image), is_image)
torch.cuda.synchronize()
torch.cuda.profiler.start()
rec = m.decode(m.encode(x, is_image), is_image)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done", float(rec.abs().mean()))
