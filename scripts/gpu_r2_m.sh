#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_track.py tests/test_gpu_tracking_class.py -q -m gpu > gpurun_out/m_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/m_tests.log
grep -n "^E  .*it [0-9]\|passed\|failed\|out of bounds\|Error\|rc=" gpurun_out/m_tests.log | cut -c1-300 | tail -12
run() {  # B G occ
  unset COMO_B200_TRACK_G COMO_B200_TRACK_OCC
  if [ "$2" != "0" ]; then export COMO_B200_TRACK_G=$2 COMO_B200_TRACK_OCC=$3; fi
  timeout 300 python bench.py --workload track640 --batch $1 --steps 20 --warmup 3 --no-e2e 1 > gpurun_out/m_sweep.json 2>gpurun_out/m_sweep.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/m_sweep.json").read().strip().splitlines()[-1])
    print("sweep B=$1 G=$2 occ=$3", round(d["value"]), "it/s kernel-frac", round(d["roofline"]["frac"],3), "ms", round(d["roofline"]["launch_ms"],4))
except Exception as e:
    print("sweep B=$1 G=$2 occ=$3 failed", e, open("gpurun_out/m_sweep.err").read()[-300:])
PY
}
run 55 8 3
run 74 6 3
run 88 5 3
run 111 4 3
run 148 3 3
run 222 2 3
run 444 1 3
export COMO_B200_TRACK_G=1 COMO_B200_TRACK_OCC=3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:track_pyr -s 3 -c 1 -f -o gpurun_out/trk_v5_b444 \
  python bench.py --workload track640 --batch 444 --steps 1 --warmup 3 --no-e2e 1 > gpurun_out/m_ncu.log 2>&1
tail -2 gpurun_out/m_ncu.log
unset COMO_B200_TRACK_G COMO_B200_TRACK_OCC
timeout 300 python bench.py --workload track640 --batch 1 --steps 20 --warmup 3 > gpurun_out/m_b1_e2e.json 2>gpurun_out/m_b1_e2e.err
python -c "
import json
d=json.loads(open('gpurun_out/m_b1_e2e.json').read().strip().splitlines()[-1]); print('B=1 e2e', d['e2e'], 'kernel ms', d['roofline']['launch_ms'])"
