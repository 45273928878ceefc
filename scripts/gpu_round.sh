mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ba.py -x -q 2>&1 | tail -15
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_ba_dmma.json 2> gpurun_out/bench_ba_dmma.err
python -c "import json; d=json.loads(open('gpurun_out/bench_ba_dmma.json').read()); print('overlap ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], d['finite'])"
COMO_B200_BA_OVERLAP=0 timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_ba_dmma0.json 2> gpurun_out/bench_ba_dmma0.err
python -c "import json; d=json.loads(open('gpurun_out/bench_ba_dmma0.json').read()); print('no-overlap ms/step', d['ms_per_step'])"
