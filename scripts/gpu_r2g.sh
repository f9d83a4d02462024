#!/bin/bash
# Round-2 GPU call G: packed-f32x2 softmax in the f16 attention core (now the default core): op + model parity, bench,
# kernel-level PEG timings, launch list and ncu of the attention core and the persistent PEG kernel.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > $O/r2g_$name.log 2>&1; echo "rc=$? $(tail -n 3 $O/r2g_$name.log | tr '\n' '|' | cut -c1-300)"; }
TMO=300 run ops_attn_h python -m pytest tests/test_gpu_f16x3.py -x -q -k "attn_spatial_h or qkv_planes"
TMO=900 OMT_TEST_VARIANTS=fast,base OMT_TEST_MATH=f16x3 run model python -m pytest tests/test_gpu_model.py -x -q -s
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
TMO=300 run ab_default $B
TMO=300 OMT_STATIC_U=1 OMT_PEG_KERNEL=5 run ab_fast $B
TMO=300 run bench_peg python scripts/bench_peg.py
TMO=300 run launches ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/r2g_launches.csv python scripts/profile_step.py f16x3
TMO=400 run ncu_attn ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attn_f16 -s 1 -c 1 -f -o $O/r2g_full_attn_f16 python scripts/profile_step.py f16x3
TMO=400 OMT_PEG_KERNEL=5 run ncu_peg5 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:peg_tile5 -s 1 -c 1 -f -o $O/r2g_full_peg5 python scripts/profile_step.py f16x3
TMO=900 run bench_cfg4 python bench.py --workload cfg4 --steps 5 --warmup 3
python scripts/launch_summary.py $O/r2g_launches.csv | head -16
cat $O/r2g_bench_peg.log
for f in ab_default ab_fast bench_cfg4; do tail -n 1 $O/r2g_$f.log | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read())
    print('$f', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'clk', d.get('clocks', {}).get('sm_mhz'), d.get('parity'))
except Exception as e:
    print('$f', 'unparsed', e)
"; done
