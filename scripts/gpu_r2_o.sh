#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for c in 0 296 148 112 96 80 64 48; do
  COMO_B200_STREAM_CTAS=$c timeout 300 python bench.py --workload ba_window --steps 20 --warmup 5 --no-e2e 1 > gpurun_out/o_ba.json 2>gpurun_out/o_ba.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/o_ba.json").read().strip().splitlines()[-1])
    print("stream ctas $c:", round(d["ms_per_step"],4), "ms/iter", round(d["value"],1), "it/s")
except Exception as e:
    print("ctas $c failed", e, open("gpurun_out/o_ba.err").read()[-300:])
PY
done
