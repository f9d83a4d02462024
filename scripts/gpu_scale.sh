#!/bin/bash
# Strong-scaling runs on one N-GPU box: python bench.py at N = 1 and torchrun at N = $@ (default 2 4 8) + the NCCL parity test.
cd "$(dirname "$0")/.." || exit 1
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
show() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], 'gpus', d['value'], 'frames/s', d['ms_per_step'], 'ms/step e2e', d['e2e']['value'], 'u8', d['e2e'].get('u8', {}).get('value'), 'gathered==single', d.get('gathered_codes_equal_single_gpu'), d['clocks'])"; }
NS=${@:-2 4 8}
timeout 300 python -m pytest tests/test_gpu_dist.py -x -q 2>&1 | tail -n 2
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | tee gpurun_out/scale_1.json | show
for n in $NS; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 20 --warmup 5 2>gpurun_out/scale_$n.err | tail -1 | tee gpurun_out/scale_$n.json | show
done
