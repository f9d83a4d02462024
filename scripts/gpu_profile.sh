#!/bin/bash
# Launch list (per-kernel device time) of one bench step.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
MATH=${1:-3xtf32}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/launches_$MATH.csv python scripts/profile_step.py $MATH > gpurun_out/ncu_step_$MATH.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_step_$MATH.log
