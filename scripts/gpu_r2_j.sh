#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_handoff.py -q -m gpu > gpurun_out/j_handoff.log 2>&1
tail -4 gpurun_out/j_handoff.log
run() {  # lib B G occ
  if [ "$1" != "default" ]; then export COMO_B200_LIB=$PWD/como_b200/var/$1; else unset COMO_B200_LIB; fi
  export COMO_B200_TRACK_G=$3 COMO_B200_TRACK_OCC=$4
  timeout 300 python bench.py --workload track640 --batch $2 --steps 10 --warmup 3 --no-e2e 1 > gpurun_out/j_sweep.json 2>gpurun_out/j_sweep.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/j_sweep.json").read().strip().splitlines()[-1])
    print("sweep $1 B=$2 G=$3 occ=$4", round(d["value"]), "it/s kernel-frac", round(d["roofline"]["frac"],3), "ms", round(d["roofline"]["launch_ms"],3))
except Exception as e:
    print("sweep $1 B=$2 G=$3 occ=$4 failed", e, open("gpurun_out/j_sweep.err").read()[-300:])
PY
}
run default 148 3 3
run default 222 2 3
run libcomo_w8.so 74 4 2
run libcomo_w8.so 148 2 2
run libcomo_w8.so 296 1 2
unset COMO_B200_LIB COMO_B200_TRACK_G COMO_B200_TRACK_OCC
