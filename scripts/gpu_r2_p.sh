#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_solve.py tests/test_gpu_ba.py -q -m gpu > gpurun_out/p_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/p_tests.log
tail -4 gpurun_out/p_tests.log
timeout 120 python scripts/chol_probe.py > gpurun_out/p_probe.log 2>&1; cat gpurun_out/p_probe.log
timeout 300 python bench.py --workload ba_window --steps 20 --warmup 5 --no-e2e 1 > gpurun_out/p_ba.json 2>gpurun_out/p_ba.err
python -c "
import json
d=json.loads(open('gpurun_out/p_ba.json').read().strip().splitlines()[-1]); print('ba', round(d['ms_per_step'],4), 'ms/iter', round(d['value'],1))"
