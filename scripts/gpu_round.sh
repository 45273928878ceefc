mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_ba_n1.json 2> gpurun_out/bench_ba_n1.err; cut -c1-200 gpurun_out/bench_ba_n1.json; python -c "import json; d=json.loads(open('gpurun_out/bench_ba_n1.json').read()); print(d['e2e']['value'], d['roofline']['frac'], d['roofline']['traffic'], d['gpu_launches'])"
