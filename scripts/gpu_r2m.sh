#!/bin/bash
# Round-2 GPU call M: statically scaled GEGLU planes as the default (single-accumulator FF2): parity at full size and in the
# three kernel sets, launch list of the final tree, bench at batch 8 / 4 / 2 (is a half batch cheaper per video?).
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > $O/r2m_$name.log 2>&1; echo "rc=$? $(tail -n 3 $O/r2m_$name.log | tr '\n' '|' | cut -c1-300)"; }
TMO=900 OMT_TEST_MATH=f16x3 run model python -m pytest tests/test_gpu_model.py -x -q
TMO=600 OMT_TEST_MATH_FULL=f16x3 run fullsize python -m pytest tests/test_gpu_fullsize.py -x -q -s
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
TMO=300 run bench_a $B
TMO=300 OMT_STATIC_U=0 run bench_u0 $B
TMO=300 run bench_b $B
TMO=300 OMT_BENCH_BATCH=4 run bench_b4 $B
TMO=300 OMT_BENCH_BATCH=2 run bench_b2 $B
TMO=300 run launches ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/fin_launches.csv python scripts/profile_step.py f16x3
TMO=300 OMT_BENCH_BATCH=1 run launches_b1 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/fin_launches_b1.csv python scripts/profile_step.py f16x3
python scripts/launch_summary.py $O/fin_launches.csv 2>/dev/null | head -16
for f in bench_a bench_u0 bench_b bench_b4 bench_b2; do tail -n 1 $O/r2m_$f.log | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read())
    print('$f', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'clk', d.get('clocks', {}).get('sm_mhz'))
except Exception as e:
    print('$f', 'unparsed', e)
"; done
