#!/bin/bash
set -x
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --profile-eager"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dw_fwd_tma_h_kernel -s 110 -c 1 -o gpurun_out/ncu_dw_fwd_tma -f $B > gpurun_out/ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pw_gemm_tc_kernel -s 243 -c 1 -o gpurun_out/ncu_pw_gemm_expand2 -f $B > gpurun_out/ncu3.log 2>&1
ls -la gpurun_out/*.ncu-rep
