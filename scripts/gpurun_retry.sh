#!/bin/bash
# gpurun with retries on "transient" (no box / pod draining: nothing charged): scripts/gpurun_retry.sh <out.txt> <timeout> <cmd>
out=$1; tmo=$2; shift 2
for i in 1 2 3 4 5 6 7 8; do
  /usr/local/graft/bin/gpurun --timeout "$tmo" -- "$@" > "$out" 2>&1
  if grep -q "status=transient\|rc=None" "$out"; then sleep 120; continue; fi
  break
done
tail -5 "$out"
