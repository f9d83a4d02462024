#!/bin/bash
export PYTHONUNBUFFERED=1
show() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', d['n_gpus'], 'gpus', d['value'], 'frames/s', d['ms_per_step'], 'ms/step e2e', d['e2e']['value'], d['clocks'])"; }
OMT_BENCH_BATCH=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | show "1gpu-B1"
OMT_BENCH_BATCH=2 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 10 --warmup 3 2>&1 | tail -1 | show "2gpu-B2"
OMT_BENCH_NO_GATHER=1 OMT_BENCH_BATCH=2 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 10 --warmup 3 2>&1 | tail -1 | show "2gpu-B2-nogather"
