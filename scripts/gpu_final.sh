#!/bin/bash
# Round-end artefacts: launch list of one step, ncu --set full of the three top kernels, the bench line.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
bash scripts/gpu_profile.sh 3xtf32
# FF1 (+GEGLU) is the 5th GEMM launch of a step (patch-embed x2, qkv, out, ff1)
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_tc2_kernel -s 4 -c 1 -f -o gpurun_out/full_gemm_ff1 python scripts/profile_step.py 3xtf32 > gpurun_out/ncu_full_gemm.log 2>&1; echo "gemm rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attn_tc3_kernel -s 1 -c 1 -f -o gpurun_out/full_attn_tc3 python scripts/profile_step.py 3xtf32 > gpurun_out/ncu_full_attn.log 2>&1; echo "attn rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:peg_tile_kernel -s 2 -c 1 -f -o gpurun_out/full_peg python scripts/profile_step.py 3xtf32 > gpurun_out/ncu_full_peg.log 2>&1; echo "peg rc=$?"
timeout 600 ncu --set full --clock-control none --profile-from-start off -k regex:vq_search_kernel -c 1 -f -o gpurun_out/full_vq python scripts/profile_step.py 3xtf32 > gpurun_out/ncu_full_vq.log 2>&1; echo "vq rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_final.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>/dev/null; tail -c 300 gpurun_out/bench_reference.json
