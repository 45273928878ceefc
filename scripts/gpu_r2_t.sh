#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python scripts/stream_stress.py 2>&1 | tail -13
timeout 600 python scripts/stream_race_check.py 2>&1 | tail -6
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/t_all_tests.log 2>&1
echo "all gpu tests rc=$?" >> gpurun_out/t_all_tests.log
tail -5 gpurun_out/t_all_tests.log
timeout 300 python bench.py --workload ba_window --steps 20 --warmup 5 --no-e2e 1 > gpurun_out/t_ba.json 2>gpurun_out/t_ba.err
python -c "
import json
d=json.loads(open('gpurun_out/t_ba.json').read().strip().splitlines()[-1]); print('ba', round(d['ms_per_step'],4), 'ms/iter', round(d['value'],1))"
timeout 300 python bench.py --workload track640 --steps 20 --warmup 3 --no-e2e 1 > gpurun_out/t_trk.json 2>gpurun_out/t_trk.err
python -c "
import json
d=json.loads(open('gpurun_out/t_trk.json').read().strip().splitlines()[-1]); print('track', round(d['value']), 'it/s frac', round(d['roofline']['frac'],3), 'ms', round(d['roofline']['launch_ms'],4))"
