#!/bin/bash
# round-2 validation: whole GPU suite, smoke, CRF micro-benchmark before/after, default bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 200 python tools/bench_crf.py 8 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_final.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "clocks")}, d["e2e"]["value"], d["sustained"]["value"])
print(d["crf"])
PY
timeout 200 python tools/bench_crf.py 8 2>&1 | tail -1
