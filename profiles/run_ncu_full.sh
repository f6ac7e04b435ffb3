#!/bin/bash
# Full ncu captures (one launch each) of the depthwise TMA kernels at the C=960, 64x64 layer.  Run under gpurun.
set -x
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --profile-eager"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dw_fwd_tma_h_kernel -s 110 -c 1 -o gpurun_out/ncu_dw_fwd_tma_v2 -f $B > gpurun_out/ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dw_wgrad_tma_h_kernel -s 52 -c 1 -o gpurun_out/ncu_dw_wgrad_tma_v2 -f $B > gpurun_out/ncu2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bn_bwd_apply_h_kernel -s 160 -c 1 -o gpurun_out/ncu_bn_bwd_apply_h -f $B > gpurun_out/ncu3.log 2>&1
ls -la gpurun_out/*.ncu-rep
