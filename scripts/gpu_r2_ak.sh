#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ba.py tests/test_gpu_edges.py -q -x 2>&1 | tail -3
timeout 300 python bench.py --workload ba_window --steps 20 --warmup 5 --no-e2e 1 > gpurun_out/ak_ba.json 2>gpurun_out/ak_ba.err
python -c "
import json
d=json.loads(open('gpurun_out/ak_ba.json').read().strip().splitlines()[-1]); print('ba', round(d['ms_per_step'],4), 'ms/iter', round(d['value'],1), 'finite', d['finite'], 'err', d['final_total_err'])"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ak_launches.csv python bench.py --workload ba_window --steps 2 --warmup 3 --no-e2e 1 > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/ak_launches.csv 2 "ba_window timed steps" > gpurun_out/ak_summary.txt; head -6 gpurun_out/ak_summary.txt
