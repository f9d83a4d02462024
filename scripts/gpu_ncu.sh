#!/bin/bash
# ncu --set full of the three top kernels inside one real step (second launch of each)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for k in gemm_tc2_kernel attn_tc_kernel peg_tile_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$k -s 2 -c 1 -f -o gpurun_out/full_$k python scripts/profile_step.py 3xtf32 > gpurun_out/ncu_full_$k.log 2>&1
  echo "$k rc=$?"
done
