#!/bin/bash
# Confirmation of the final tree: the driver's own sequence (pytest -m gpu, smoke, bench, reference arm) on one box.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > $O/cfm_$name.log 2>&1; echo "rc=$? $(tail -n 3 $O/cfm_$name.log | tr '\n' '|' | cut -c1-300)"; }
TMO=1800 run pytest_gpu python -m pytest tests/ -x -q -m gpu
TMO=300 run smoke python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
TMO=600 run bench_cfg3 python bench.py --gpus 1 --steps 20 --warmup 5
TMO=900 run reference python bench.py --impl reference --gpus 1 --steps 3 --warmup 1
TMO=300 run bench_cfg3_b python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline
for f in bench_cfg3 reference bench_cfg3_b; do tail -n 1 $O/cfm_$f.log | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read())
    print('$f', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e'].get('u8', {}).get('value'), 'clk', d.get('clocks', {}).get('sm_mhz'), 'roof', d.get('roofline', {}).get('frac'), d.get('parity'), d.get('cpu_baseline', {}).get('sample', '')[:160])
except Exception as e:
    print('$f', 'unparsed', e)
"; done
