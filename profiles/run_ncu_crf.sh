#!/bin/bash
# Full ncu captures of the dense-CRF mean-field kernels (tools/bench_crf.py 8 --once; second call = warm).
# usage: bash profiles/run_ncu_crf.sh <tag>
tag=$1
mkdir -p gpurun_out
B="python tools/bench_crf.py 8 --once"
# launch indices inside one call: the build takes ~45 launches; skip into the mean-field loop of the SECOND call
for spec in "crf_splat_kernel:26:splat_gauss" "crf_splat_kernel:27:splat_bilat" "crf_slice2_kernel:12:slice2" "crf_blur_kernel:120:blur"; do
  IFS=: read -r k s o <<< "$spec"
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c 1 -o gpurun_out/ncu_${tag}_crf_$o -f $B > gpurun_out/ncu_${tag}_crf_$o.log 2>&1
done
ls -la gpurun_out/*crf*.ncu-rep
