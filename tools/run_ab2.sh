#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
bash tools/run_ab.sh DLB_BN_STAGE_KB 24 12 16 32 48
timeout 300 ncu --set full --clock-control none --import-source on -c 1 -f -k regex:resize_softmax_ce -s 2 -o gpurun_out/ncu_r02_resize_softmax_ce \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-crf --profile-eager > gpurun_out/ncu_ce.log 2>&1
ls -la gpurun_out/ncu_r02_resize_softmax_ce.ncu-rep
