#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_track.py tests/test_gpu_tracking_class.py -q -m gpu > gpurun_out/e_track_tests.log 2>&1
echo "track tests rc=$?" >> gpurun_out/e_track_tests.log
grep -n "AssertionError: (\|passed\|failed\|Error" gpurun_out/e_track_tests.log | tail -12
for cfg in "1 0 0" "74 6 3" "111 4 3" "148 3 3" "222 2 3"; do
  set -- $cfg
  unset COMO_B200_TRACK_G COMO_B200_TRACK_OCC
  if [ "$2" != "0" ]; then export COMO_B200_TRACK_G=$2 COMO_B200_TRACK_OCC=$3; fi
  timeout 300 python bench.py --workload track640 --batch $1 --steps 10 --warmup 3 --no-e2e 1 > gpurun_out/e_sweep_b$1_g$2_o$3.json 2>gpurun_out/e_sweep_b$1.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/e_sweep_b$1_g$2_o$3.json").read().strip().splitlines()[-1])
    print("sweep B=$1 G=$2 occ=$3", round(d["value"]), "it/s step-frac", round(d["roofline"]["alg_bytes_per_launch"]/d["ms_per_step"]/1e6/d["roofline"]["peak"],3), "kernel-frac", round(d["roofline"]["frac"],3), "ms", round(d["ms_per_step"],3), round(d["roofline"]["launch_ms"],3))
except Exception as e:
    print("sweep B=$1 G=$2 occ=$3 failed", e)
PY
done
unset COMO_B200_TRACK_G COMO_B200_TRACK_OCC
timeout 900 python -m pytest tests/test_gpu_ba.py -q -m gpu -x > gpurun_out/e_ba_tests.log 2>&1
tail -5 gpurun_out/e_ba_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/e_bench_all.json 2> gpurun_out/e_bench_all.err
tail -c 3000 gpurun_out/e_bench_all.json; tail -5 gpurun_out/e_bench_all.err
