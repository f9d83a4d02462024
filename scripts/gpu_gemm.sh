#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "linear" --timeout 120 -p no:cacheprovider -x 2>&1 | tail -15
timeout 300 python scripts/profile_gemm.py time 2>&1 | tee gpurun_out/gemm_times.txt
