#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_track.py tests/test_gpu_tracking_class.py -q -m gpu > gpurun_out/f_track_tests.log 2>&1
echo "track tests rc=$?" >> gpurun_out/f_track_tests.log
grep -n "AssertionError: (\|passed\|failed\|Error" gpurun_out/f_track_tests.log | tail -12
export COMO_B200_TRACK_G=3 COMO_B200_TRACK_OCC=3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:track_pyr -s 3 -c 1 -f -o gpurun_out/trk_v4_b148 \
  python bench.py --workload track640 --batch 148 --steps 1 --warmup 3 --no-e2e 1 > gpurun_out/f_ncu.log 2>&1
tail -2 gpurun_out/f_ncu.log
unset COMO_B200_TRACK_G COMO_B200_TRACK_OCC
for G in 16 32 64 96 148; do
  COMO_B200_TRACK_G=$G COMO_B200_TRACK_OCC=1 timeout 300 python bench.py --workload track640 --batch 1 --steps 20 --warmup 3 --no-e2e 1 > gpurun_out/f_b1_g$G.json 2>/dev/null
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/f_b1_g$G.json").read().strip().splitlines()[-1])
    print("B=1 G=$G occ=1 ms/frame", round(d["roofline"]["launch_ms"],4), "step", round(d["ms_per_step"],4))
except Exception as e:
    print("B=1 G=$G failed", e)
PY
done
