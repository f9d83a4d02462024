#!/bin/bash
# Round-2 final 1-GPU call: the driver's own sequence (pytest -m gpu, smoke, bench + reference arm) and every number under
# profiles/: bench lines of all BASELINE.json configurations, math-mode A/B, launch lists, ncu --set full of the hot kernels.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > $O/fin_$name.log 2>&1; echo "rc=$? $(tail -n 3 $O/fin_$name.log | tr '\n' '|' | cut -c1-300)"; }
TMO=1800 run pytest_gpu python -m pytest tests/ -x -q -m gpu
TMO=300 run smoke python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
TMO=600 run bench_cfg3 python bench.py --gpus 1 --steps 20 --warmup 5
TMO=900 run reference python bench.py --impl reference --gpus 1 --steps 5 --warmup 1
TMO=300 run bench_cfg3_3xtf32 python bench.py --math 3xtf32 --steps 20 --warmup 5 --no-cpu-baseline
TMO=300 OMT_BENCH_BATCH=1 run bench_cfg3_b1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline
TMO=600 run bench_cfg2 python bench.py --workload cfg2 --steps 10 --warmup 3
TMO=900 run bench_cfg4 python bench.py --workload cfg4 --steps 5 --warmup 3
TMO=600 run bench_cfg5 python bench.py --workload cfg5 --steps 10 --warmup 3
TMO=300 run gemm_shapes python scripts/bench_gemm_shapes.py
TMO=300 run launches ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/fin_launches.csv python scripts/profile_step.py f16x3
TMO=300 OMT_BENCH_BATCH=1 run launches_b1 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/fin_launches_b1.csv python scripts/profile_step.py f16x3
TMO=400 run ncu_attn ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attn_f16 -s 1 -c 1 -f -o $O/fin_full_attn_f16 python scripts/profile_step.py f16x3
TMO=400 run ncu_ff1 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_f16_kernel -s 12 -c 1 -f -o $O/fin_full_gemm_ff1 python scripts/profile_step.py f16x3
TMO=400 run ncu_peg ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:peg_tile4 -s 1 -c 1 -f -o $O/fin_full_peg4 python scripts/profile_step.py f16x3
TMO=400 run ncu_vq ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:vq_fused -c 1 -f -o $O/fin_full_vq python scripts/profile_step.py f16x3
TMO=400 run ncu_ln ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:layernorm_kernel -s 3 -c 1 -f -o $O/fin_full_ln python scripts/profile_step.py f16x3
python scripts/launch_summary.py $O/fin_launches.csv 2>/dev/null | head -12
for f in bench_cfg3 reference bench_cfg3_3xtf32 bench_cfg3_b1 bench_cfg2 bench_cfg4 bench_cfg5; do tail -n 1 $O/fin_$f.log | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read())
    print('$f', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'clk', d.get('clocks', {}).get('sm_mhz'), 'roof', d.get('roofline', {}).get('frac'), d.get('parity'))
except Exception as e:
    print('$f', 'unparsed', e)
"; done
