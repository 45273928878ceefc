#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/k_all_tests.log 2>&1
echo "all gpu tests rc=$?" >> gpurun_out/k_all_tests.log
tail -6 gpurun_out/k_all_tests.log
run() {  # lib B G occ rcap
  if [ "$1" != "default" ]; then export COMO_B200_LIB=$PWD/como_b200/var/$1; else unset COMO_B200_LIB; fi
  export COMO_B200_TRACK_G=$3 COMO_B200_TRACK_OCC=$4
  if [ "$5" != "" ]; then export COMO_B200_TRACK_RCAP=$5; else unset COMO_B200_TRACK_RCAP; fi
  timeout 300 python bench.py --workload track640 --batch $2 --steps 10 --warmup 3 --no-e2e 1 > gpurun_out/k_sweep.json 2>gpurun_out/k_sweep.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/k_sweep.json").read().strip().splitlines()[-1])
    print("sweep $1 B=$2 G=$3 occ=$4 rcap=$5", round(d["value"]), "it/s kernel-frac", round(d["roofline"]["frac"],3), "ms", round(d["roofline"]["launch_ms"],3))
except Exception as e:
    print("sweep $1 B=$2 G=$3 occ=$4 failed", e, open("gpurun_out/k_sweep.err").read()[-300:])
PY
}
run libcomo_s2.so 148 3 3 0
run libcomo_s2.so 222 2 3 0
run libcomo_s2.so 148 3 3
run default 148 3 3 0
unset COMO_B200_LIB COMO_B200_TRACK_RCAP
export COMO_B200_TRACK_G=64 COMO_B200_TRACK_OCC=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:track_pyr -s 3 -c 1 -f -o gpurun_out/trk_v4_b1 \
  python bench.py --workload track640 --batch 1 --steps 1 --warmup 3 --no-e2e 1 > gpurun_out/k_ncu.log 2>&1
tail -2 gpurun_out/k_ncu.log
