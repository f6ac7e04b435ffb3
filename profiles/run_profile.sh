#!/bin/bash
# Run on the GPU box (under gpurun): launch list of one training step + full capture of the top kernels.
# Outputs land in gpurun_out/ ; summaries are copied into profiles/ by hand afterwards.
set -x
mkdir -p gpurun_out
# every launch with its device time (cold-cache, serialised: compare SHARES, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --profile-eager > gpurun_out/ncu_bench.log 2>&1
