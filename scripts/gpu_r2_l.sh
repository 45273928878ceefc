#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_track.py tests/test_gpu_tracking_class.py tests/test_gpu_handoff.py -q -m gpu > gpurun_out/l_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/l_tests.log
grep -n "^E  .*it [0-9]\|passed\|failed\|out of bounds\|Error\|rc=" gpurun_out/l_tests.log | cut -c1-300 | tail -20
run() {  # B G occ
  unset COMO_B200_TRACK_G COMO_B200_TRACK_OCC
  if [ "$2" != "0" ]; then export COMO_B200_TRACK_G=$2 COMO_B200_TRACK_OCC=$3; fi
  timeout 300 python bench.py --workload track640 --batch $1 --steps 20 --warmup 3 --no-e2e 1 > gpurun_out/l_sweep.json 2>gpurun_out/l_sweep.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/l_sweep.json").read().strip().splitlines()[-1])
    print("sweep B=$1 G=$2 occ=$3", round(d["value"]), "it/s kernel-frac", round(d["roofline"]["frac"],3), "ms", round(d["roofline"]["launch_ms"],4))
except Exception as e:
    print("sweep B=$1 G=$2 occ=$3 failed", e, open("gpurun_out/l_sweep.err").read()[-300:])
PY
}
run 1 0 0
run 1 32 1
run 1 64 1
run 1 148 1
run 1 296 2
run 74 0 0
run 148 3 3
run 222 2 3
run 296 1 2
run 444 1 3
