#!/bin/bash
# full single-GPU evidence: all GPU tests, smoke, default bench, launch list
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/z_pytest_gpu.log 2>&1
echo "pytest -m gpu rc=$?" >> gpurun_out/z_pytest_gpu.log
tail -3 gpurun_out/z_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/z_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/z_smoke.log; tail -2 gpurun_out/z_smoke.log
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/z_bench_all_n1.json 2>gpurun_out/z_bench_all_n1.err
echo "bench rc=$?"; tail -c 400 gpurun_out/z_bench_all_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/z_bench_all_n1.json').read().strip().splitlines()[-1])
print('headline', round(d['value'],2), d['unit'], 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1))
print('roofline', round(d['roofline']['frac'],3), 'warp', round(d['roofline_warp']['frac'],3), 'step', round(d['roofline_step']['frac'],3), 'cpu', d['cpu_baseline']['value'])
for k,v in (d.get('secondary') or {}).items():
    print(' ', k, round(v.get('value',0),2), v.get('unit'), 'ms', round(v.get('ms_per_step',0),4), 'frac', (v.get('roofline') or {}).get('frac'), 'e2e', (v.get('e2e') or {}).get('value'))
PY
