#!/bin/bash
# quick loop: op tests + model parity (3xtf32) + launch list + short bench
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu --timeout 300 -p no:cacheprovider -x 2>&1 | tail -5
OMT_TEST_MATH=3xtf32 timeout 600 python -m pytest tests/test_gpu_model.py -q -m gpu --timeout 600 -p no:cacheprovider -x 2>&1 | tail -5
bash scripts/gpu_profile.sh 3xtf32
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-400
