#!/bin/bash
# Full ncu captures of the pointwise GEMM alone (tools/bench_gemm.py --ncu: second launch of one shape, L2 flushed).
# usage: bash profiles/run_ncu_gemm.sh <tag> 160x960 960x160 ...
tag=$1; shift
mkdir -p gpurun_out
for shp in "$@"; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:pw_gemm_tc -s 1 -c 1 \
    -o gpurun_out/ncu_${tag}_gemm_${shp} -f python tools/bench_gemm.py --ncu $shp > gpurun_out/ncu_${tag}_gemm_${shp}.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
