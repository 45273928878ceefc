#!/bin/bash
# final single-GPU validation of round 2: tests, smoke, both bench arms, launch list
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/v_gpu.txt
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/v_pytest_gpu.log 2>&1
echo "pytest -m gpu rc=$?" >> gpurun_out/v_pytest_gpu.log
tail -4 gpurun_out/v_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/v_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/v_smoke.log; tail -3 gpurun_out/v_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/v_bench_all_n1.json 2>gpurun_out/v_bench_all_n1.err
echo "bench rc=$?"; tail -c 600 gpurun_out/v_bench_all_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/v_bench_all_n1.json').read().strip().splitlines()[-1])
print('headline', d['metric'], round(d['value'],2), d['unit'], 'ms', round(d['ms_per_step'],4), 'e2e', d.get('e2e'))
print('roofline', d.get('roofline')); print('cpu', d.get('cpu_baseline')); print('clocks', d.get('clocks'))
for k,v in (d.get('secondary') or {}).items():
    print(' ', k, json.dumps(v)[:420])
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/v_bench_ref_cpu.json 2>gpurun_out/v_bench_ref_cpu.err
echo "ref cpu rc=$?"; tail -1 gpurun_out/v_bench_ref_cpu.json | cut -c1-500
timeout 600 python bench.py --impl reference --device cuda --steps 3 --warmup 1 > gpurun_out/v_bench_ref_cuda.json 2>gpurun_out/v_bench_ref_cuda.err
echo "ref cuda rc=$?"; tail -1 gpurun_out/v_bench_ref_cuda.json | cut -c1-500
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/v_launches_all.csv python bench.py --steps 2 --warmup 3 > gpurun_out/v_ncu_bench.log 2>&1
echo "ncu rc=$?"
python scripts/launch_summary.py gpurun_out/v_launches_all.csv 2 'python bench.py --steps 2 --warmup 3 (timed steps only, cudaProfilerStart/Stop)' > gpurun_out/v_launches_summary.txt 2>&1; head -40 gpurun_out/v_launches_summary.txt
