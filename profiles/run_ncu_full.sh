#!/bin/bash
# full ncu captures (with source) of the kernels that dominate the training step, C=960 layer of block 14/16
set -x
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --profile-eager"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dw_fwd_tiled_h_kernel -s 110 -c 1 -o gpurun_out/ncu_dw_fwd_h -f $B > gpurun_out/ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dw_wgrad_tiled_h_kernel -s 51 -c 1 -o gpurun_out/ncu_dw_wgrad_h -f $B > gpurun_out/ncu2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pw_gemm_tc_kernel -s 243 -c 1 -o gpurun_out/ncu_pw_gemm_expand -f $B > gpurun_out/ncu3.log 2>&1
ls -la gpurun_out/*.ncu-rep
