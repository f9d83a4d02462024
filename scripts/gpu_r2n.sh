#!/bin/bash
# Round-2 GPU call N: PEG v4 without the out-of-volume planes in shared memory (69 KB tile, three CTAs per SM): op tests
# (bit-identical to v3), kernel timings, model parity, bench, launch list.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > $O/r2n_$name.log 2>&1; echo "rc=$? $(tail -n 3 $O/r2n_$name.log | tr '\n' '|' | cut -c1-300)"; }
TMO=300 run ops_peg python -m pytest tests/test_gpu_ops.py -x -q -k "peg"
TMO=300 run bench_peg python scripts/bench_peg.py
TMO=900 OMT_TEST_MATH=f16x3 OMT_TEST_VARIANTS=default run model python -m pytest tests/test_gpu_model.py -x -q
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
TMO=300 run bench_a $B
TMO=300 run bench_b $B
TMO=300 run launches ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/fin_launches.csv python scripts/profile_step.py f16x3
TMO=300 OMT_BENCH_BATCH=1 run launches_b1 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/fin_launches_b1.csv python scripts/profile_step.py f16x3
TMO=400 run ncu_peg ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:peg_tile4 -s 1 -c 1 -f -o $O/fin_full_peg4 python scripts/profile_step.py f16x3
cat $O/r2n_bench_peg.log
python scripts/launch_summary.py $O/fin_launches.csv 2>/dev/null | head -9
for f in bench_a bench_b; do tail -n 1 $O/r2n_$f.log | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read())
    print('$f', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'clk', d.get('clocks', {}).get('sm_mhz'))
except Exception as e:
    print('$f', 'unparsed', e)
"; done
