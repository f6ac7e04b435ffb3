#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -q -x -k "resize_softmax or small_gemm" 2>&1 | tail -4
timeout 200 python bench.py --no-crf --no-bf16 --no-cpu-baseline > gpurun_out/bench_ce2.json 2>/dev/null
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_ce2.json").read().strip().splitlines()[-1])
f = d["roofline"]["per_family_ms"]
print("ms/step", round(d["ms_per_step"], 4), "sustained", round(d["sustained"]["ms_per_step"], 4), {k: f[k] for k in ("resize_softmax_ce", "small_gemm")})
PY
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4
