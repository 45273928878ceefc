#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python scripts/chol_probe.py 2>&1 | tail -12
timeout 600 python -m pytest tests/test_gpu_solve.py tests/test_gpu_ba.py -q -x 2>&1 | tail -4
timeout 300 python bench.py --workload ba_window --steps 20 --warmup 5 --no-e2e 1 > gpurun_out/aa_ba.json 2>gpurun_out/aa_ba.err
python -c "
import json
d=json.loads(open('gpurun_out/aa_ba.json').read().strip().splitlines()[-1]); print('ba', round(d['ms_per_step'],4), 'ms/iter', round(d['value'],1))"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/aa_launches.csv python bench.py --workload ba_window --steps 2 --warmup 3 --no-e2e 1 > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/aa_launches.csv 2 "ba_window timed steps" | head -8
