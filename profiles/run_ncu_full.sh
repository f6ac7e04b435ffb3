#!/bin/bash
# full ncu captures of the three kernels that dominate the training step (C=960 layer of block 14), 1 launch each
set -x
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --profile-eager"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dw_fwd_tiled_kernel -s 110 -c 1 -o gpurun_out/ncu_dw_fwd -f $B > gpurun_out/ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bn_bwd_reduce_kernel -s 166 -c 1 -o gpurun_out/ncu_bn_bwd_reduce -f $B > gpurun_out/ncu2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pw_gemm_tc_kernel -s 243 -c 1 -o gpurun_out/ncu_pw_gemm_expand -f $B > gpurun_out/ncu3.log 2>&1
ls -la gpurun_out/*.ncu-rep
