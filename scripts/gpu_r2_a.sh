#!/bin/bash
# round 2, call A: new tracking kernel -- parity first, then the whole GPU suite, then batch sweeps
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/a_gpu.txt
timeout 900 python -m pytest tests/test_gpu_track.py -x -q -m gpu > gpurun_out/a_track_tests.log 2>&1
echo "track tests rc=$?" >> gpurun_out/a_track_tests.log
tail -5 gpurun_out/a_track_tests.log
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/a_all_tests.log 2>&1
echo "all tests rc=$?" >> gpurun_out/a_all_tests.log
tail -5 gpurun_out/a_all_tests.log
for B in 1 16 37 74; do
  timeout 300 python bench.py --workload track640 --batch $B --steps 10 --warmup 3 > gpurun_out/a_track_b$B.json 2> gpurun_out/a_track_b$B.err
  tail -c 1500 gpurun_out/a_track_b$B.json
done
for cfg in "37 4 2" "37 8 2" "18 8 1" "18 16 2" "74 2 1" "74 4 2" "148 2 2" "148 1 1"; do
  set -- $cfg
  COMO_B200_TRACK_G=$2 COMO_B200_TRACK_OCC=$3 timeout 300 python bench.py --workload track640 --batch $1 --steps 10 --warmup 3 --no-e2e 1 > gpurun_out/a_sweep_b$1_g$2_o$3.json 2>/dev/null
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/a_sweep_b$1_g$2_o$3.json").read().strip().splitlines()[-1])
    print("sweep B=$1 G=$2 occ=$3", round(d["value"]), "it/s frac", round(d["roofline"]["frac"],3), "ms", round(d["ms_per_step"],3))
except Exception as e:
    print("sweep B=$1 G=$2 occ=$3 failed", e)
PY
done
