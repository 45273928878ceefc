mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_solve.py tests/test_gpu_ba.py -x -q 2>&1 | tail -3
timeout 120 python scripts/chol_timeline.py 2848 2>&1 | tail -16
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_ba_pair.json 2> gpurun_out/bench_ba_pair.err
python -c "import json; d=json.loads(open('gpurun_out/bench_ba_pair.json').read()); print('ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], d['finite'], d['final_total_err'])"
