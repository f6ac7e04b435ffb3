#!/bin/bash
# Run on the GPU box (under gpurun): tests, the bench line, the per-call log, the ncu launch list of one training
# step and full captures of the top kernels.  Outputs land in gpurun_out/ ; digests are copied into profiles/ afterwards
# (tools/ncu_digest.py, tools/summarize_launches.py).
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q > gpurun_out/tests_full.log 2>&1; tail -2 gpurun_out/tests_full.log
DLB_CALL_LOG=gpurun_out/calls_r01.jsonl timeout 400 python bench.py > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; cut -c1-250 gpurun_out/bench_r01.json
timeout 200 python tools/bench_gemm.py > gpurun_out/gemm_bench_r01.jsonl 2>> gpurun_out/bench_r01.err
timeout 200 python tools/bench_layer.py > gpurun_out/layer_bench_r01.jsonl 2>> gpurun_out/bench_r01.err; cat gpurun_out/layer_bench_r01.jsonl
# every launch with its device time (cold-cache, serialised: compare SHARES, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --profile-eager > gpurun_out/ncu_bench.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -c 1 -f"
timeout 200 $NCU -k regex:pw_gemm_tc -s 1 -o gpurun_out/ncu_r01_pw_gemm_expand python tools/bench_gemm.py --ncu 160x960 > gpurun_out/ncu_a.log 2>&1
timeout 200 $NCU -k regex:bn_stream -s 2 -o gpurun_out/ncu_r01_bn_bwd_reduce_c960 python tools/bench_layer.py --ncu bn_bwd > gpurun_out/ncu_b.log 2>&1
timeout 200 $NCU -k regex:bn_stream -s 3 -o gpurun_out/ncu_r01_bn_bwd_apply_c960 python tools/bench_layer.py --ncu bn_bwd > gpurun_out/ncu_c.log 2>&1
timeout 200 $NCU -k regex:dw_fwd_tma_h -s 1 -o gpurun_out/ncu_r01_dw_fwd_c960 python tools/bench_layer.py --ncu dw_conv_fwd > gpurun_out/ncu_d.log 2>&1
timeout 200 $NCU -k regex:dw_wgrad_tma_h -s 1 -o gpurun_out/ncu_r01_dw_wgrad_c960 python tools/bench_layer.py --ncu dw_conv_bwd > gpurun_out/ncu_e.log 2>&1
timeout 200 $NCU -k regex:pw_wgrad_tc -s 1 -o gpurun_out/ncu_r01_pw_wgrad python tools/bench_layer.py --ncu pw_wgrad > gpurun_out/ncu_f.log 2>&1
ls -la gpurun_out/*.ncu-rep
