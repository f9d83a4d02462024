#!/bin/bash
# Round-2 GPU call J: attention core with Q in shared memory (SS-form S MMAs; tensor-memory read traffic 2048 -> 1280 cycles per
# key tile): op + model parity in both kernel shapes, same-box A/B, launch lists, ncu.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > $O/r2j_$name.log 2>&1; echo "rc=$? $(tail -n 3 $O/r2j_$name.log | tr '\n' '|' | cut -c1-300)"; }
TMO=300 run ops_attn_h python -m pytest tests/test_gpu_f16x3.py -x -q -k "attn_spatial_h"
TMO=900 OMT_TEST_MATH=f16x3 OMT_TEST_VARIANTS=default,fast run model python -m pytest tests/test_gpu_model.py -x -q
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
TMO=300 run ab_default $B
TMO=300 OMT_ATTN_CTAS=2 run ab_ctas2 $B
TMO=300 run ab_default2 $B
TMO=300 OMT_ATTN_CTAS=2 run ab_ctas2_2 $B
TMO=300 run launches ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/r2j_launches.csv python scripts/profile_step.py f16x3
TMO=300 OMT_ATTN_CTAS=2 run launches2 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/r2j_launches2.csv python scripts/profile_step.py f16x3
TMO=400 OMT_ATTN_CTAS=2 run ncu_attn ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attn_f16 -s 1 -c 1 -f -o $O/r2j_full_attn_f16 python scripts/profile_step.py f16x3
python scripts/launch_summary.py $O/r2j_launches.csv 2>/dev/null | grep attn
python scripts/launch_summary.py $O/r2j_launches2.csv 2>/dev/null | grep attn
for f in ab_default ab_ctas2 ab_default2 ab_ctas2_2; do tail -n 1 $O/r2j_$f.log | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read())
    print('$f', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'clk', d.get('clocks', {}).get('sm_mhz'))
except Exception as e:
    print('$f', 'unparsed', e)
"; done
