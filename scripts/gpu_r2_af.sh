#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_track.py tests/test_gpu_tracking_class.py -q -x > gpurun_out/af_tests.log 2>&1
echo "track tests rc=$?"; tail -12 gpurun_out/af_tests.log
for b in 592 444; do
timeout 300 python bench.py --batch $b --workload track640 --steps 10 --warmup 3 --no-e2e 1 > gpurun_out/af_b$b.json 2>gpurun_out/af_b$b.err
python -c "
import json
d=json.loads(open('gpurun_out/af_b$b.json').read().strip().splitlines()[-1]); print('B=$b', round(d['value']), 'it/s frac', round(d['roofline']['frac'],4), 'ms', round(d['roofline']['launch_ms'],4))"
done
