mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ba.py tests/test_gpu_edges.py -x -q 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench_ba_n1.json 2> gpurun_out/bench_ba_n1.err; python -c "import json; d=json.loads(open('gpurun_out/bench_ba_n1.json').read()); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['finite'])"; tail -2 gpurun_out/bench_ba_n1.err
