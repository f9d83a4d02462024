#!/bin/bash
# Round-2 GPU call O: LayerNorm with 8 consecutive columns per lane (16-byte plane stores): op tests, model parity, launch list, bench.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > $O/r2o_$name.log 2>&1; echo "rc=$? $(tail -n 3 $O/r2o_$name.log | tr '\n' '|' | cut -c1-300)"; }
TMO=600 run ops python -m pytest tests/test_gpu_ops.py tests/test_gpu_f16x3.py -x -q
TMO=900 OMT_TEST_MATH=f16x3 OMT_TEST_VARIANTS=default,base run model python -m pytest tests/test_gpu_model.py -x -q
TMO=300 run launches ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/r2o_launches.csv python scripts/profile_step.py f16x3
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
TMO=300 run bench_a $B
TMO=300 run bench_b $B
python scripts/launch_summary.py $O/r2o_launches.csv 2>/dev/null | grep -E "layernorm|total"
for f in bench_a bench_b; do tail -n 1 $O/r2o_$f.log | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read())
    print('$f', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'clk', d.get('clocks', {}).get('sm_mhz'))
except Exception as e:
    print('$f', 'unparsed', e)
"; done
