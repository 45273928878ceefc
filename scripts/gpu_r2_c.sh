#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_track.py tests/test_gpu_tracking_class.py -q -m gpu > gpurun_out/c_track_tests.log 2>&1
echo "track tests rc=$?" >> gpurun_out/c_track_tests.log
tail -15 gpurun_out/c_track_tests.log
for cfg in "1 0 0" "37 8 2" "74 4 2" "148 2 2"; do
  set -- $cfg
  if [ "$2" != "0" ]; then export COMO_B200_TRACK_G=$2 COMO_B200_TRACK_OCC=$3; fi
  timeout 300 python bench.py --workload track640 --batch $1 --steps 10 --warmup 3 --no-e2e 1 > gpurun_out/c_sweep_b$1_g$2_o$3.json 2>/dev/null
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/c_sweep_b$1_g$2_o$3.json").read().strip().splitlines()[-1])
    print("sweep B=$1 G=$2 occ=$3", round(d["value"]), "it/s frac", round(d["roofline"]["frac"],3), "ms", round(d["ms_per_step"],3))
except Exception as e:
    print("sweep B=$1 G=$2 occ=$3 failed", e)
PY
done
