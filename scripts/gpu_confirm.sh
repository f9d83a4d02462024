#!/bin/bash
# Confirmation of the final tree: the driver's own sequence (pytest -m gpu, smoke, bench) on one box.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > $O/cfm_$name.log 2>&1; echo "rc=$? $(tail -n 3 $O/cfm_$name.log | tr '\n' '|' | cut -c1-300)"; }
TMO=1800 run pytest_gpu python -m pytest tests/ -x -q -m gpu
TMO=300 run smoke python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
TMO=600 run bench_cfg3 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline
tail -n 1 $O/cfm_bench_cfg3.log | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('bench_cfg3', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e'].get('u8', {}).get('value'), 'clk', d.get('clocks', {}).get('sm_mhz'), 'roof', d.get('roofline', {}).get('frac'))
"
