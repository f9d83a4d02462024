#!/bin/bash
# Round-2 GPU call E: fused VQ kernel, uint8 e2e, f16 attention core bring-up (op tests in their own process), benches of
# every BASELINE.json configuration and the CPU reference arm; launch lists and ncu of the kernels as shipped.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > $O/r2e_$name.log 2>&1; echo "rc=$? $(tail -n 3 $O/r2e_$name.log | tr '\n' '|' | cut -c1-300)"; }
TMO=600 run ops python -m pytest tests/test_gpu_ops.py tests/test_gpu_f16x3.py tests/test_gpu_consumers.py -x -q -k "not attn_spatial_h and not qkv_planes"
TMO=300 run ops_qkvplanes python -m pytest tests/test_gpu_f16x3.py -x -q -k "qkv_planes"
TMO=300 run ops_attn_h python -m pytest tests/test_gpu_f16x3.py -x -q -k "attn_spatial_h"
TMO=900 OMT_TEST_VARIANTS=base run model python -m pytest tests/test_gpu_model.py -x -q
TMO=600 OMT_TEST_VARIANTS=upeg OMT_TEST_MATH=f16x3 run model_u_peg python -m pytest tests/test_gpu_model.py -x -q -s -k "golden"
TMO=900 OMT_TEST_VARIANTS=fast OMT_TEST_MATH=f16x3 run model_fast python -m pytest tests/test_gpu_model.py -x -q -s
TMO=600 run bench_cfg3 python bench.py --steps 20 --warmup 5
TMO=400 OMT_STATIC_U=1 run bench_cfg3_u python bench.py --steps 10 --warmup 3 --no-cpu-baseline
TMO=400 OMT_PEG_KERNEL=5 run bench_cfg3_peg5 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
TMO=400 OMT_ATTN_F16=1 run bench_cfg3_attn_h python bench.py --steps 10 --warmup 3 --no-cpu-baseline
TMO=600 OMT_ATTN_F16=1 OMT_STATIC_U=1 OMT_PEG_KERNEL=5 run bench_cfg3_fast python bench.py --steps 20 --warmup 5
TMO=600 run bench_cfg2 python bench.py --workload cfg2 --steps 10 --warmup 3
TMO=900 run bench_cfg4 python bench.py --workload cfg4 --steps 5 --warmup 3
TMO=900 OMT_ATTN_F16=1 run bench_cfg4_attn_h python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu-baseline
TMO=600 run bench_cfg5 python bench.py --workload cfg5 --steps 10 --warmup 3
TMO=300 run launches ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/r2e_launches.csv python scripts/profile_step.py f16x3
TMO=300 OMT_ATTN_F16=1 run launches_attn_h ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/r2e_launches_attn_h.csv python scripts/profile_step.py f16x3
TMO=300 OMT_BENCH_BATCH=1 run launches_b1 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/r2e_launches_b1.csv python scripts/profile_step.py f16x3
TMO=400 run ncu_ff1 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_f16_kernel -s 12 -c 1 -f -o $O/r2e_full_gemm_ff1 python scripts/profile_step.py f16x3
TMO=400 run ncu_vq ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:vq_fused -c 1 -f -o $O/r2e_full_vq python scripts/profile_step.py f16x3
TMO=400 OMT_ATTN_F16=1 run ncu_attn_h ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attn_f16 -s 1 -c 1 -f -o $O/r2e_full_attn_f16 python scripts/profile_step.py f16x3
TMO=900 run reference python bench.py --impl reference --steps 3 --warmup 1
python scripts/launch_summary.py $O/r2e_launches.csv | head -24
for f in bench_cfg3 bench_cfg3_u bench_cfg3_peg5 bench_cfg3_attn_h bench_cfg3_fast bench_cfg2 bench_cfg4 bench_cfg4_attn_h bench_cfg5 reference; do tail -n 1 $O/r2e_$f.log | cut -c1-260; done
