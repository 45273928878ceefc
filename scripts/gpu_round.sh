mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ba.py tests/test_gpu_solve.py -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_ba_tiled2.json 2> gpurun_out/bench_ba_tiled2.err
python -c "import json; d=json.loads(open('gpurun_out/bench_ba_tiled2.json').read()); print('tiled ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], d['finite'], d['final_total_err'])"
