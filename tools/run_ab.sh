#!/bin/bash
# A/B of an environment switch on the bench line: tools/run_ab.sh VAR v1 v2 ...
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
var=$1; shift
for v in "$@"; do
  env $var=$v timeout 200 python bench.py --no-crf --no-bf16 --no-cpu-baseline > gpurun_out/ab_$v.json 2>/dev/null
  python - "$var=$v" gpurun_out/ab_$v.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
f = d["roofline"]["per_family_ms"]
print(sys.argv[1], "ms/step", round(d["ms_per_step"], 4), "sustained", round(d["sustained"]["ms_per_step"], 4), "launches", d["gpu_launches"] // d["steps"],
      {k: f[k] for k in ("bn_bwd", "bn_act_apply")})
PY
done
