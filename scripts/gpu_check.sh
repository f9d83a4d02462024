#!/bin/bash
# One gpurun call: parity tests (fp32 path, then tcgen05 path in separate processes), smoke, short bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
export PYTHONUNBUFFERED=1
T="timeout 900"
$T python -m pytest tests/test_gpu_ops.py -q -m gpu -k "not 3xtf32 and not tf32" --timeout 300 -p no:cacheprovider > gpurun_out/ops_fp32.log 2>&1; echo "ops_fp32 rc=$?"
$T python -m pytest tests/test_gpu_ops.py -q -m gpu -k "3xtf32 or tf32" --timeout 300 -p no:cacheprovider > gpurun_out/ops_tc.log 2>&1; echo "ops_tc rc=$?"
OMT_TEST_MATH=fp32 $T python -m pytest tests/test_gpu_model.py -q -m gpu --timeout 600 -s -p no:cacheprovider > gpurun_out/model_fp32.log 2>&1; echo "model_fp32 rc=$?"
OMT_TEST_MATH=3xtf32 $T python -m pytest tests/test_gpu_model.py -q -m gpu --timeout 600 -s -p no:cacheprovider -k "golden and not vae or intermediate or forward" > gpurun_out/model_tc.log 2>&1; echo "model_tc rc=$?"
$T python bench.py --steps 3 --warmup 3 --math fp32 > gpurun_out/bench_fp32.log 2>&1; echo "bench_fp32 rc=$?"
$T python bench.py --steps 5 --warmup 3 --math 3xtf32 > gpurun_out/bench_3xtf32.log 2>&1; echo "bench_3xtf32 rc=$?"
tail -n 30 gpurun_out/ops_fp32.log gpurun_out/ops_tc.log gpurun_out/model_fp32.log gpurun_out/model_tc.log
tail -n 3 gpurun_out/bench_fp32.log gpurun_out/bench_3xtf32.log
