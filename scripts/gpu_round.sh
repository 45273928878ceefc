set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ba.py -x -q 2>&1 | tail -3
for cfg in "1 s" "0 s"; do
  set -- $cfg
  COMO_B200_BA_OVERLAP=$1 COMO_B200_PREDICTOR=$2 timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_ba_$1$2.json 2> gpurun_out/bench_ba_$1$2.err
  python -c "import json,sys; d=json.loads(open('gpurun_out/bench_ba_$1$2.json').read()); print('$1$2', d['ms_per_step'], d['e2e']['value'], d['roofline']['launch_ms'])"
done
