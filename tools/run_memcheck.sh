#!/bin/bash
# compute-sanitizer memcheck over the consumer-side BatchNorm finalize paths (bn_stream, project GEMM A transform, depthwise prologue)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_ops_gpu.py -q -x \
  -k "consumer_side and (16384-96 or 8192-960 or 16-256 or 16384-576 or 1024-32 or 64-96 or 32-960 or 16-64-1-36 or argument)" > gpurun_out/memcheck_fin.log 2>&1
echo "memcheck rc=$?"; grep -c "Invalid\|out of bounds" gpurun_out/memcheck_fin.log; tail -6 gpurun_out/memcheck_fin.log
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/memcheck_smoke.log 2>&1
echo "smoke memcheck rc=$?"; tail -3 gpurun_out/memcheck_smoke.log
