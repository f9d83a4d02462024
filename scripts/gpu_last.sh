#!/bin/bash
# Last gpurun call of the round, most important first: full GPU suite + smoke on the final build, same-box A/B of the
# remote-arrive scope (step and per-GEMM), one-video-per-GPU step, launch lists (B=8 and B=1) of the final build.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 420 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider 2>&1 | tail -n 6
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 4
for sc in 1 0; do
  OMT_TC_ARRIVE_CTA=$sc timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/last_scope$sc.json 2> gpurun_out/last_scope$sc.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/last_scope$sc.json").read().strip().splitlines()[-1])
    print("arrive_cta=$sc B=8", d["value"], "frames/s", d["ms_per_step"], "ms/step; FF1", d["roofline"]["achieved"], "TF/s", d["roofline"]["ms_per_launch"], "ms; e2e", d["e2e"]["value"], d["e2e"].get("host_link"))
except Exception as e:
    print("arrive_cta=$sc failed", e); print(open("gpurun_out/last_scope$sc.err").read()[-1500:])
PY
done
timeout 120 python scripts/bench_gemm_scope.py 2>&1 | tail -n 9 | tee gpurun_out/last_gemm_scope.txt
OMT_BENCH_BATCH=1 timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/last_b1.json 2> gpurun_out/last_b1.err
python -c "
import json
d = json.loads(open('gpurun_out/last_b1.json').read().strip().splitlines()[-1]); print('B=1', d['value'], 'frames/s', d['ms_per_step'], 'ms/step')"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/last_launches_b8.csv python scripts/profile_step.py 3xtf32 > gpurun_out/last_ncu_b8.log 2>&1; echo "ncu b8 rc=$?"
OMT_BENCH_BATCH=1 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/last_launches_b1.csv python scripts/profile_step.py 3xtf32 > gpurun_out/last_ncu_b1.log 2>&1; echo "ncu b1 rc=$?"
