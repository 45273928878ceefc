#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for mode in 1 0; do
echo "== schedule $mode"
COMO_B200_CHOL_SCHEDULE=$mode timeout 300 python -m pytest tests/test_gpu_solve.py -q -x 2>&1 | tail -3
COMO_B200_CHOL_SCHEDULE=$mode timeout 300 python bench.py --workload ba_window --steps 20 --warmup 5 --no-e2e 1 > gpurun_out/ah_ba_$mode.json 2>gpurun_out/ah_ba_$mode.err
python -c "
import json
d=json.loads(open('gpurun_out/ah_ba_$mode.json').read().strip().splitlines()[-1]); print('ba', round(d['ms_per_step'],4), 'ms/iter', round(d['value'],1), 'finite', d['finite'], 'err', d['final_total_err'])"
done
COMO_B200_CHOL_SCHEDULE=1 timeout 300 python -m pytest tests/test_gpu_ba.py -q -x 2>&1 | tail -3
COMO_B200_CHOL_SCHEDULE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ah_launches.csv python bench.py --workload ba_window --steps 2 --warmup 3 --no-e2e 1 > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/ah_launches.csv 2 "ba_window timed steps" > gpurun_out/ah_summary.txt; head -6 gpurun_out/ah_summary.txt
