mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kfinit.py -q 2>&1 | tail -5
timeout 900 python bench.py --workload kf_init --steps 10 --warmup 3 > gpurun_out/bench_kfinit.json 2> gpurun_out/bench_kfinit.err; cat gpurun_out/bench_kfinit.json; tail -5 gpurun_out/bench_kfinit.err
