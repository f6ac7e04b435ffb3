#!/bin/bash
# Round-2 evidence run (under gpurun, 1 GPU): smoke, the bench line with its per-call log, the ncu launch list of one
# eager training step, full ncu captures of the kernels changed this round.  Outputs land in gpurun_out/; digests are
# written into profiles/ afterwards with tools/ncu_digest.py / tools/summarize_launches.py.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/smoke_r02.log 2>&1; tail -2 gpurun_out/smoke_r02.log
DLB_CALL_LOG=gpurun_out/calls_r02.jsonl timeout 400 python bench.py > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err; cut -c1-300 gpurun_out/bench_r02.json
timeout 200 python tools/bench_gemm.py > gpurun_out/gemm_bench_r02.jsonl 2>> gpurun_out/bench_r02.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-crf --profile-eager > gpurun_out/ncu_bench_r02.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -c 1 -f"
timeout 200 $NCU -k regex:pw_gemm_tc -s 1 -o gpurun_out/ncu_r02_pw_gemm_expand python tools/bench_gemm.py --ncu 160x960 > gpurun_out/ncu_a.log 2>&1
timeout 200 $NCU -k regex:pw_gemm_tc -s 1 -o gpurun_out/ncu_r02_pw_gemm_16x96 python tools/bench_gemm.py --ncu 16x96 > gpurun_out/ncu_b.log 2>&1
timeout 200 $NCU -k regex:pw_gemm_tc -s 1 -o gpurun_out/ncu_r02_pw_gemm_project_xform python tools/bench_gemm.py --ncu --xform 960x160 > gpurun_out/ncu_c.log 2>&1
timeout 200 $NCU -k regex:pw_wgrad_tc -s 1 -o gpurun_out/ncu_r02_pw_wgrad_xform python tools/bench_layer.py --ncu pw_wgrad > gpurun_out/ncu_d.log 2>&1
B="python tools/bench_crf.py 8 --once"
for spec in "crf_splat_kernel:26:splat_gauss" "crf_splat_kernel:27:splat_bilat" "crf_slice2_kernel:12:slice2"; do
  IFS=: read -r k s o <<< "$spec"
  timeout 300 $NCU -k regex:$k -s $s -o gpurun_out/ncu_r02_crf_$o $B > gpurun_out/ncu_crf_$o.log 2>&1
done
timeout 200 $NCU -k regex:sepconv_fused -s 20 -o gpurun_out/ncu_r02_sepconv_mbconv python tools/xception_once.py float16 mobilenetv2 16 > gpurun_out/ncu_e.log 2>&1
ls -la gpurun_out/*r02*.ncu-rep
