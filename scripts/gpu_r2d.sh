#!/bin/bash
# Round-2 GPU call D: row-scaled single-accumulator GEMM form (LayerNorm / patch-gather fed shapes): op + model tests,
# GEMM shape timings of all three forms, bench.  Output: gpurun_out/r2d_*.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > $O/r2d_$name.log 2>&1; echo "rc=$? $(tail -n 3 $O/r2d_$name.log | tr '\n' '|' | cut -c1-300)"; }
TMO=420 run ops python -m pytest tests/test_gpu_f16x3.py tests/test_gpu_ops.py -x -q
TMO=900 OMT_TEST_MATH=f16x3 run model python -m pytest tests/test_gpu_model.py tests/test_gpu_consumers.py -x -q
TMO=300 run gemm_shapes python scripts/bench_gemm_shapes.py 40960 5120
TMO=600 run bench python bench.py --steps 10 --warmup 3
TMO=300 OMT_BENCH_BATCH=1 run bench_b1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
TMO=600 run fullsize python -m pytest tests/test_gpu_fullsize.py -x -q -s
grep TFLOP $O/r2d_gemm_shapes.log
for f in bench bench_b1; do tail -n 1 $O/r2d_$f.log | cut -c1-300; done
