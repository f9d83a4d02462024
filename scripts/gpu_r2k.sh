#!/bin/bash
# Round-2 GPU call K: same-box, same-process A/B of three builds of the f16 attention core (Q as a TMEM operand / Q in shared
# memory / + P_hi.[V_hi|V_lo] as one N = 128 MMA), then parity of the working tree.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > $O/r2k_$name.log 2>&1; echo "rc=$? $(tail -n 3 $O/r2k_$name.log | tr '\n' '|' | cut -c1-300)"; }
TMO=300 run ops_attn_h python -m pytest tests/test_gpu_f16x3.py -x -q -k "attn_spatial_h"
TMO=300 run bench_attn python scripts/bench_attn.py scripts/ubench/lib_002e53c.so scripts/ubench/lib_b639ea7.so omnitokenizer_b200/libomnitok_b200.so
TMO=900 OMT_TEST_MATH=f16x3 OMT_TEST_VARIANTS=default,fast run model python -m pytest tests/test_gpu_model.py -x -q
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
TMO=300 run ab_default $B
TMO=300 OMT_ATTN_CTAS=2 run ab_ctas2 $B
cat $O/r2k_bench_attn.log
for f in ab_default ab_ctas2; do tail -n 1 $O/r2k_$f.log | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read())
    print('$f', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'clk', d.get('clocks', {}).get('sm_mhz'))
except Exception as e:
    print('$f', 'unparsed', e)
"; done
