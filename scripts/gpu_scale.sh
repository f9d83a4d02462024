#!/bin/bash
export PYTHONUNBUFFERED=1
N=${1:-2}
for b in 8 1; do OMT_BENCH_BATCH=$b timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('1gpu B=$b', d['value'], 'frames/s', d['ms_per_step'], 'ms/step e2e', d['e2e']['value'])"; done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 2>&1 | tail -3 | cut -c1-600
