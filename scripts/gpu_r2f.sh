#!/bin/bash
# Round-2 GPU call F: ALU issue-rate microbenchmark, the fused-VQ op test, same-step A/B of the three optional kernels
# (f16 attention core, PEG v5, static U planes) and the worker-process reference arm.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > $O/r2f_$name.log 2>&1; echo "rc=$? $(tail -n 3 $O/r2f_$name.log | tr '\n' '|' | cut -c1-300)"; }
TMO=120 run ubench scripts/ubench/alu
TMO=300 run ops_vq python -m pytest tests/test_gpu_ops.py -x -q -k "vq"
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
TMO=300 run ab_base1 $B
TMO=300 OMT_ATTN_F16=1 run ab_attn1 $B
TMO=300 OMT_PEG_KERNEL=5 run ab_peg1 $B
TMO=300 OMT_STATIC_U=1 run ab_u1 $B
TMO=300 OMT_ATTN_F16=1 OMT_STATIC_U=1 OMT_PEG_KERNEL=5 run ab_fast1 $B
TMO=300 run ab_base2 $B
TMO=300 OMT_ATTN_F16=1 OMT_PEG_KERNEL=5 run ab_attn_peg2 $B
TMO=300 OMT_ATTN_F16=1 OMT_STATIC_U=1 OMT_PEG_KERNEL=5 run ab_fast2 $B
TMO=900 run reference python bench.py --impl reference --steps 3 --warmup 1
TMO=900 run bench_cfg4 python bench.py --workload cfg4 --steps 5 --warmup 3
cat $O/r2f_ubench.log
for f in ab_base1 ab_attn1 ab_peg1 ab_u1 ab_fast1 ab_base2 ab_attn_peg2 ab_fast2 reference bench_cfg4; do tail -n 1 $O/r2f_$f.log | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read())
    print('$f', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'clk', d.get('clocks', {}).get('sm_mhz'), d.get('parity'), d.get('cpu_baseline', {}).get('cores'))
except Exception as e:
    print('$f', 'unparsed', e)
"; done
