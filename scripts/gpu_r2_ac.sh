#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_depthcov.py tests/test_gpu_kfinit.py -q -x 2>&1 | tail -5
timeout 600 python bench.py --workload kf_init --steps 20 --warmup 5 > gpurun_out/ac_kfinit.json 2>gpurun_out/ac_kfinit.err
python -c "
import json
d=json.loads(open('gpurun_out/ac_kfinit.json').read().strip().splitlines()[-1]); print('kf_init', round(d['ms_per_step'],3), 'ms', round(d['value'],1), 'KF/s e2e', round(d['e2e']['value'],1))"
tail -c 300 gpurun_out/ac_kfinit.err
