#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_track.py tests/test_gpu_tracking_class.py -q -m gpu > gpurun_out/d_track_tests.log 2>&1
echo "track tests rc=$?" >> gpurun_out/d_track_tests.log
grep -n "AssertionError: (\|passed\|failed" gpurun_out/d_track_tests.log | tail -12
for cfg in "1 0 0" "37 12 3" "74 6 3" "74 4 2" "111 4 3" "148 3 3" "148 2 2" "222 2 3"; do
  set -- $cfg
  unset COMO_B200_TRACK_G COMO_B200_TRACK_OCC
  if [ "$2" != "0" ]; then export COMO_B200_TRACK_G=$2 COMO_B200_TRACK_OCC=$3; fi
  timeout 300 python bench.py --workload track640 --batch $1 --steps 10 --warmup 3 --no-e2e 1 > gpurun_out/d_sweep_b$1_g$2_o$3.json 2>/dev/null
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/d_sweep_b$1_g$2_o$3.json").read().strip().splitlines()[-1])
    print("sweep B=$1 G=$2 occ=$3", round(d["value"]), "it/s frac", round(d["roofline"]["frac"],3), "ms", round(d["ms_per_step"],3))
except Exception as e:
    print("sweep B=$1 G=$2 occ=$3 failed", e)
PY
done
export COMO_B200_TRACK_G=6 COMO_B200_TRACK_OCC=3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:track_pyr -s 3 -c 1 -f -o gpurun_out/trk_v3_b74 \
  python bench.py --workload track640 --batch 74 --steps 1 --warmup 3 --no-e2e 1 > gpurun_out/d_ncu.log 2>&1
tail -2 gpurun_out/d_ncu.log
