mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_ba_n1.json 2> gpurun_out/bench_ba_n1.err; cut -c1-260 gpurun_out/bench_ba_n1.json; python -c "import json; d=json.loads(open('gpurun_out/bench_ba_n1.json').read()); print(d['e2e'], d['roofline'])"
timeout 600 ncu --set full --clock-control none --profile-from-start off -k regex:predictor_stream_kernel -c 1 -o gpurun_out/predictor_stream python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_pred.log 2>&1; tail -1 gpurun_out/ncu_pred.log
