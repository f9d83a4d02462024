#!/bin/bash
# Round-2 GPU call H: attention core with the O accumulator kept in tensor memory (lazy running maximum), group-minimum VQ
# kernel; op + model + full-size parity, bench A/B of the PEG kernels, launch list, ncu of attention and VQ.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > $O/r2h_$name.log 2>&1; echo "rc=$? $(tail -n 3 $O/r2h_$name.log | tr '\n' '|' | cut -c1-300)"; }
TMO=300 run ops_vq python -m pytest tests/test_gpu_ops.py -x -q -k "vq"
TMO=300 run ops_attn_h python -m pytest tests/test_gpu_f16x3.py -x -q -k "attn_spatial_h or qkv_planes"
TMO=900 OMT_TEST_VARIANTS=fast,base OMT_TEST_MATH=f16x3 run model python -m pytest tests/test_gpu_model.py -x -q -s
TMO=900 OMT_TEST_MATH_FULL=f16x3 run fullsize python -m pytest tests/test_gpu_fullsize.py -x -q -s
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
TMO=300 run ab_default $B
TMO=300 OMT_PEG_KERNEL=4 run ab_peg4 $B
TMO=300 OMT_PEG_KERNEL=4 OMT_STATIC_U=1 run ab_peg4_u $B
TMO=300 run ab_default2 $B
TMO=300 OMT_BENCH_BATCH=1 run ab_b1 $B
TMO=300 OMT_PEG_KERNEL=4 run launches ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/r2h_launches.csv python scripts/profile_step.py f16x3
TMO=400 run ncu_attn ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attn_f16 -s 1 -c 1 -f -o $O/r2h_full_attn_f16 python scripts/profile_step.py f16x3
TMO=400 run ncu_vq ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:vq_fused -c 1 -f -o $O/r2h_full_vq python scripts/profile_step.py f16x3
TMO=600 run bench_cfg4 python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu-baseline
python scripts/launch_summary.py $O/r2h_launches.csv 2>/dev/null | head -18
for f in ab_default ab_peg4 ab_peg4_u ab_default2 ab_b1 bench_cfg4; do tail -n 1 $O/r2h_$f.log | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read())
    print('$f', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'clk', d.get('clocks', {}).get('sm_mhz'), d.get('vq_lookup', {}).get('us'))
except Exception as e:
    print('$f', 'unparsed', e)
"; done
