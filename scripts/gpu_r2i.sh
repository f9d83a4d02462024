#!/bin/bash
# Round-2 GPU call I: PEG v4 as the default, the two-CTAs-per-SM shape of the attention core; full op suites, model parity in
# all three kernel sets, same-box A/B.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > $O/r2i_$name.log 2>&1; echo "rc=$? $(tail -n 3 $O/r2i_$name.log | tr '\n' '|' | cut -c1-300)"; }
TMO=300 run ops_attn_h python -m pytest tests/test_gpu_f16x3.py -x -q -k "attn_spatial_h"
TMO=900 run ops python -m pytest tests/test_gpu_ops.py tests/test_gpu_f16x3.py tests/test_gpu_consumers.py -x -q -k "not attn_spatial_h"
TMO=900 OMT_TEST_MATH=f16x3 run model python -m pytest tests/test_gpu_model.py -x -q
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
TMO=300 run ab_default $B
TMO=300 OMT_ATTN_CTAS=2 run ab_ctas2 $B
TMO=300 OMT_ATTN_CTAS=2 OMT_STATIC_U=1 run ab_ctas2_u $B
TMO=300 OMT_STATIC_U=1 run ab_u $B
TMO=300 run ab_default2 $B
TMO=300 OMT_ATTN_CTAS=2 run ab_ctas2_2 $B
TMO=300 OMT_ATTN_CTAS=2 run launches ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/r2i_launches.csv python scripts/profile_step.py f16x3
TMO=400 OMT_ATTN_CTAS=2 run ncu_attn ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attn_f16 -s 1 -c 1 -f -o $O/r2i_full_attn_f16 python scripts/profile_step.py f16x3
TMO=600 OMT_ATTN_CTAS=2 run bench_cfg4 python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu-baseline
python scripts/launch_summary.py $O/r2i_launches.csv 2>/dev/null | head -10
for f in ab_default ab_ctas2 ab_ctas2_u ab_u ab_default2 ab_ctas2_2 bench_cfg4; do tail -n 1 $O/r2i_$f.log | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read())
    print('$f', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'clk', d.get('clocks', {}).get('sm_mhz'), d.get('vq_lookup', {}).get('us'))
except Exception as e:
    print('$f', 'unparsed', e)
"; done
