#!/bin/bash
# One short gpurun call: PEG v4 correctness + timing, then same-box A/B of the step (peg kernel, PDL at one video per GPU).
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T="timeout 300"
$T python -m pytest tests/test_gpu_ops.py -q -m gpu -k "peg" --timeout 200 -p no:cacheprovider 2>&1 | tail -n 15
$T python scripts/bench_peg.py 2>&1 | tail -n 8
for pk in 3 4; do
  OMT_PEG_KERNEL=$pk $T python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/exp_peg$pk.json 2> gpurun_out/exp_peg$pk.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/exp_peg$pk.json").read().strip().splitlines()[-1])
    print("peg_kernel=$pk B=8", d["value"], "frames/s", d["ms_per_step"], "ms/step e2e", d["e2e"]["value"], d["e2e"].get("host_link"), "parity", d.get("parity"))
except Exception as e:
    print("peg_kernel=$pk failed", e); print(open("gpurun_out/exp_peg$pk.err").read()[-1500:])
PY
done
for pdl in 0 1; do
  OMT_PDL=$pdl OMT_BENCH_BATCH=1 $T python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/exp_pdl$pdl.json 2> gpurun_out/exp_pdl$pdl.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/exp_pdl$pdl.json").read().strip().splitlines()[-1])
    print("pdl=$pdl B=1", d["value"], "frames/s", d["ms_per_step"], "ms/step")
except Exception as e:
    print("pdl=$pdl failed", e); print(open("gpurun_out/exp_pdl$pdl.err").read()[-1500:])
PY
done
