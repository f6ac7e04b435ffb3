#!/bin/bash
# closing validation: whole GPU suite, smoke, default bench line
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2
DLB_CALL_LOG=gpurun_out/calls_r02.jsonl timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_final.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "clocks")}, d["e2e"]["value"], d["sustained"]["value"])
print(d["roofline"]["per_family_ms"])
print(d["crf"]["ms_per_img"], d["crf"]["ms_per_call"])
PY
