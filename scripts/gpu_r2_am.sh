#!/bin/bash
# reduced final evidence: all GPU tests, default bench line, launch list of the timed steps
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/am_pytest_gpu.log 2>&1
echo "pytest -m gpu rc=$?" >> gpurun_out/am_pytest_gpu.log
tail -4 gpurun_out/am_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/am_bench_all_n1.json 2>gpurun_out/am_bench_all_n1.err
echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/am_bench_all_n1.json').read().strip().splitlines()[-1])
print('headline', round(d['value'],2), d['unit'], 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), 'roof', round(d['roofline']['frac'],3), 'warp', round(d['roofline_warp']['frac'],3), 'cpu', round(d['cpu_baseline']['value'],3))
for k,v in (d.get('secondary') or {}).items():
    print(' ', k, round(v.get('value',0),2), v.get('unit'), 'ms', round(v.get('ms_per_step',0),4), 'frac', (v.get('roofline') or {}).get('frac'), 'e2e', (v.get('e2e') or {}).get('value'), v.get('error'))
PY
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/am_launches_all.csv python bench.py --workload ba_window --steps 2 --warmup 3 > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/am_launches_all.csv 2 'python bench.py --workload ba_window --steps 2 --warmup 3 (timed steps only, cudaProfilerStart/Stop)' > gpurun_out/am_launches_summary.txt 2>&1; head -8 gpurun_out/am_launches_summary.txt | cut -c1-110
