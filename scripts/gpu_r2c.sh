#!/bin/bash
# Round-2 GPU call C: full GPU suite with f16x3 as the default math (incl. consumers / uint8 epilogue), launch list of one
# step, ncu --set full of the f16x3 GEMM shapes and the other top kernels.  Output: gpurun_out/r2c_*.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > $O/r2c_$name.log 2>&1; echo "rc=$? $(tail -n 3 $O/r2c_$name.log | tr '\n' '|' | cut -c1-300)"; }
TMO=1800 run tests python -m pytest tests -m gpu -x -q
TMO=300 run smoke python -c "import __graft_entry__ as g; g.smoke()"
TMO=300 run launches ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/r2c_launches.csv python scripts/profile_step.py f16x3
TMO=300 OMT_BENCH_BATCH=1 run launches_b1 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/r2c_launches_b1.csv python scripts/profile_step.py f16x3
for s in ff1 out qkv ff2; do
  TMO=400 run ncu_$s ncu --set full --clock-control none --import-source on -k regex:gemm_f16 -s 2 -c 1 -f -o $O/r2c_full_gemm_$s python scripts/profile_gemm.py $s
done
TMO=400 run ncu_out128 ncu --set full --clock-control none --import-source on -k regex:gemm_f16 -s 2 -c 1 -f -o $O/r2c_full_gemm_out128 python scripts/profile_gemm.py out 40960 128
for k in attn_tc3_kernel peg_tile_kernel layernorm_kernel attn_temporal_kernel attn_flash_kernel; do
  TMO=400 run ncu_$k ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$k -s 1 -c 1 -f -o $O/r2c_full_$k python scripts/profile_step.py f16x3
done
TMO=600 run bench python bench.py --steps 10 --warmup 3
python scripts/launch_summary.py $O/r2c_launches.csv | head -30
tail -n 1 $O/r2c_bench.log | cut -c1-300
