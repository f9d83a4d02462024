#!/bin/bash
# Round-2 GPU call A: first hardware contact of the f16x3 path (op tests per operand scheme in separate processes,
# model-level parity, GEMM shape timings, bench A/B against 3xTF32).  Everything lands in gpurun_out/r2a_*.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/r2a_gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > $O/r2a_$name.log 2>&1; echo "rc=$? $(tail -n 3 $O/r2a_$name.log | tr '\n' '|' | cut -c1-400)"; }
TMO=420 OMT_TEST_SCHEMES=1 run ops_s1 python -m pytest tests/test_gpu_f16x3.py -x -q
TMO=420 OMT_TEST_SCHEMES=2 run ops_s2 python -m pytest tests/test_gpu_f16x3.py -x -q -k "not multiwave"
TMO=300 run ops_old python -m pytest tests/test_gpu_ops.py -x -q
TMO=600 OMT_TEST_MATH=f16x3 run model_f16 python -m pytest tests/test_gpu_model.py -x -q -s
TMO=300 run gemm_shapes python scripts/bench_gemm_shapes.py 40960 5120
TMO=300 BOTH_SCHEMES=1 run gemm_shapes_s2 python scripts/bench_gemm_shapes.py 40960
TMO=400 run bench_3xtf32 python bench.py --math 3xtf32 --steps 10 --warmup 3 --no-cpu-baseline
TMO=600 run bench_f16x3 python bench.py --math f16x3 --steps 10 --warmup 3
TMO=300 OMT_BENCH_BATCH=1 run bench_f16x3_b1 python bench.py --math f16x3 --steps 10 --warmup 3 --no-cpu-baseline
TMO=300 OMT_BENCH_BATCH=1 run bench_3xtf32_b1 python bench.py --math 3xtf32 --steps 10 --warmup 3 --no-cpu-baseline
grep -h "frames/s\|TFLOP" $O/r2a_gemm_shapes.log | head -40
for f in bench_3xtf32 bench_f16x3 bench_f16x3_b1 bench_3xtf32_b1; do tail -n 1 $O/r2a_$f.log | cut -c1-700; done
