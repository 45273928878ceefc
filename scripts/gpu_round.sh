mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_ba_n1.json 2> gpurun_out/bench_ba_n1.err; python -c "import json; d=json.loads(open('gpurun_out/bench_ba_n1.json').read()); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['finite'], d['clocks']['reasons'])"
