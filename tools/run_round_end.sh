#!/bin/bash
# What the driver runs at round end, in one gpurun call (2 GPUs): smoke, both bench arms at N=1 and N=2.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; cut -c1-400 gpurun_out/bench_n1.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_n1.json 2> gpurun_out/bench_ref_n1.err; cut -c1-600 gpurun_out/bench_ref_n1.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; cut -c1-500 gpurun_out/bench_n2.json; tail -3 gpurun_out/bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; cut -c1-300 gpurun_out/bench_ref_n2.json
