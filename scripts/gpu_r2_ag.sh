#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_track.py -q -x 2>&1 | tail -3
for b in 592 444; do
timeout 300 python bench.py --batch $b --workload track640 --steps 10 --warmup 3 --no-e2e 1 > gpurun_out/ag_b$b.json 2>gpurun_out/ag_b$b.err
python -c "
import json
d=json.loads(open('gpurun_out/ag_b$b.json').read().strip().splitlines()[-1]); print('B=$b', round(d['value']), 'it/s frac', round(d['roofline']['frac'],4), 'ms', round(d['roofline']['launch_ms'],4))"
done
timeout 600 python bench.py --workload track640 --steps 20 --warmup 5 > gpurun_out/ag_trk.json 2>gpurun_out/ag_trk.err
python -c "
import json
d=json.loads(open('gpurun_out/ag_trk.json').read().strip().splitlines()[-1]); print('standalone with e2e', round(d['value']), 'frac', round(d['roofline']['frac'],4), 'ms', round(d['roofline']['launch_ms'],4), 'step ms', round(d['ms_per_step'],4), 'e2e', d['e2e'])"
nvidia-smi --query-gpu=temperature.gpu,power.draw,clocks.sm,clocks.mem --format=csv
