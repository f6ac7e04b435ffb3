#!/bin/bash
# Full ncu captures (one launch each) of selected kernels inside one eager training step.  Run under gpurun.
# usage: bash profiles/run_ncu_full.sh "<kernel regex>:<skip>:<out name>" ...
set -x
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --profile-eager"
for spec in "$@"; do
  IFS=: read -r k s o <<< "$spec"
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c 1 -o gpurun_out/$o -f $B > gpurun_out/$o.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
