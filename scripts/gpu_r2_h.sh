#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_track.py tests/test_gpu_tracking_class.py -q -m gpu > gpurun_out/h_track_tests.log 2>&1
echo "track tests rc=$?" >> gpurun_out/h_track_tests.log
grep -n "^E  .*it [0-9]\|passed\|failed\|out of bounds\|Error" gpurun_out/h_track_tests.log | cut -c1-330 | tail -40
for cfg in "74 6 3" "148 3 3" "222 2 3" "444 1 3" "296 1 2"; do
  set -- $cfg
  export COMO_B200_TRACK_G=$2 COMO_B200_TRACK_OCC=$3
  timeout 300 python bench.py --workload track640 --batch $1 --steps 10 --warmup 3 --no-e2e 1 > gpurun_out/h_sweep_b$1_g$2_o$3.json 2>gpurun_out/h_sweep_b$1.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/h_sweep_b$1_g$2_o$3.json").read().strip().splitlines()[-1])
    print("sweep B=$1 G=$2 occ=$3", round(d["value"]), "it/s kernel-frac", round(d["roofline"]["frac"],3), "ms", round(d["ms_per_step"],3), round(d["roofline"]["launch_ms"],3))
except Exception as e:
    print("sweep B=$1 G=$2 occ=$3 failed", e)
PY
done
