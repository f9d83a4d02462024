#!/bin/bash
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
show() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], 'gpus', d['value'], 'frames/s', d['ms_per_step'], 'ms/step e2e', d['e2e']['value'], d['clocks'])"; }
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | tee gpurun_out/scale_1.json | show
for n in 8 4; do timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 10 --warmup 3 2>/dev/null | tail -1 | tee gpurun_out/scale_$n.json | show; done
