#!/bin/bash
# Round-2 GPU call B: f16x3 path with fp16 hi/lo planes (two accumulators), tile widths 256 (single TMEM buffer, register
# drain) and 128 (double buffer): op tests, model parity, GEMM shape timings, bench A/B.  Output: gpurun_out/r2b_*.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > $O/r2b_$name.log 2>&1; echo "rc=$? $(tail -n 3 $O/r2b_$name.log | tr '\n' '|' | cut -c1-300)"; }
TMO=420 OMT_TEST_F16_BN=256 run ops_256 python -m pytest tests/test_gpu_f16x3.py -x -q
TMO=420 OMT_TEST_F16_BN=128 run ops_128 python -m pytest tests/test_gpu_f16x3.py -x -q -k "not producers"
TMO=600 OMT_TEST_MATH=f16x3 run model_f16 python -m pytest tests/test_gpu_model.py -x -q -s
TMO=300 run gemm_shapes python scripts/bench_gemm_shapes.py 40960 5120
TMO=600 run bench_f16x3 python bench.py --math f16x3 --steps 10 --warmup 3
TMO=400 OMT_F16_BN=128 run bench_f16x3_bn128 python bench.py --math f16x3 --steps 10 --warmup 3 --no-cpu-baseline
TMO=300 OMT_BENCH_BATCH=1 run bench_f16x3_b1 python bench.py --math f16x3 --steps 10 --warmup 3 --no-cpu-baseline
TMO=300 OMT_BENCH_BATCH=1 OMT_F16_BN=128 run bench_f16x3_b1_bn128 python bench.py --math f16x3 --steps 10 --warmup 3 --no-cpu-baseline
cat $O/r2b_gemm_shapes.log | grep TFLOP
for f in bench_f16x3 bench_f16x3_bn128 bench_f16x3_b1 bench_f16x3_b1_bn128; do tail -n 1 $O/r2b_$f.log | cut -c1-400; done
