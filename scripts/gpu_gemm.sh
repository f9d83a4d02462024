#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python scripts/profile_gemm.py time 2>&1 | tee gpurun_out/gemm_times.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 1 -c 1 -f -o gpurun_out/gemm_ff1 python scripts/profile_gemm.py ncu > gpurun_out/ncu_gemm.log 2>&1
echo "ncu rc=$?"
