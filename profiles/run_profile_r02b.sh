#!/bin/bash
# Round-2 closing evidence run (under gpurun, 1 GPU) for the step as shipped (consumer-side BatchNorm finalize, 309
# launches): the GPU suite four times over (run-to-run spread of the atomics-ordered tests), the bench line with its
# per-call log, the ncu launch list of one eager training step, full ncu captures of the dominant family (bn_bwd).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for i in 1 2 3 4; do
  timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/suite_$i.log; tail -2 gpurun_out/suite_$i.log
done
DLB_CALL_LOG=gpurun_out/calls_r02.jsonl timeout 400 python bench.py > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err; cut -c1-260 gpurun_out/bench_r02.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-crf --profile-eager > gpurun_out/ncu_bench_r02.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -c 1 -f"
timeout 200 $NCU -k regex:bn_stream -s 2 -o gpurun_out/ncu_r02_bn_bwd_reduce_c960 python tools/bench_layer.py --ncu bn_bwd > gpurun_out/ncu_f.log 2>&1
timeout 200 $NCU -k regex:bn_stream -s 3 -o gpurun_out/ncu_r02_bn_bwd_apply_c960 python tools/bench_layer.py --ncu bn_bwd > gpurun_out/ncu_g.log 2>&1
ls -la gpurun_out/ncu_r02_bn_bwd*.ncu-rep
