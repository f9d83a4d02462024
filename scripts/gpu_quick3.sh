#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu --timeout 300 -p no:cacheprovider -x 2>&1 | tail -3
OMT_TEST_MATH=3xtf32,fp32 timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu --timeout 600 -p no:cacheprovider -x 2>&1 | tail -3
GEMM_M=5120 timeout 300 python scripts/profile_gemm.py time 2>&1 | grep "bn=2"
timeout 300 python scripts/profile_gemm.py time 2>&1 | grep "bn=2"
for b in 8 1; do OMT_BENCH_BATCH=$b timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('B=$b', d['value'], 'frames/s', d['ms_per_step'], 'ms/step e2e', d['e2e']['value'], 'launches', d['gpu_launches'])"; done
