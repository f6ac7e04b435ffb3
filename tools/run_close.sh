#!/bin/bash
# closing evidence for the shipped build: default bench line with its per-call log, ncu launch list of one eager step
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
DLB_CALL_LOG=gpurun_out/calls_r02.jsonl timeout 400 python bench.py > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err; cut -c1-200 gpurun_out/bench_r02.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-crf --profile-eager > gpurun_out/ncu_bench_r02.log 2>&1
wc -l gpurun_out/launches_r02.csv
