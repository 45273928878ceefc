#!/bin/bash
# multi-GPU evidence: N = number of GPUs of the box (gpurun --gpus N)
N=${1:-2}
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29513 scripts/dist_ba_check.py > gpurun_out/n${N}_dist_check.log 2>&1; echo "dist check rc=$?"; tail -4 gpurun_out/n${N}_dist_check.log
timeout 1500 $TR --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/n${N}_bench_all.json 2> gpurun_out/n${N}_bench_all.err; echo "bench rc=$?"
python - $N <<'PY'
import json,sys
N=sys.argv[1]
d=json.loads(open(f'gpurun_out/n{N}_bench_all.json').read().strip().splitlines()[-1])
print('headline', round(d['value'],2), d['unit'], 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), d['scaling'], d['n_gpus'])
for k,v in (d.get('secondary') or {}).items():
    print(' ', k, round(v.get('value',0),2), v.get('unit'), 'ms', round(v.get('ms_per_step',0),4), 'frac', (v.get('roofline') or {}).get('frac'), v.get('scaling'))
PY
tail -c 300 gpurun_out/n${N}_bench_all.err
timeout 600 $TR --master-port 29512 bench.py --gpus $N --impl reference --steps 2 --warmup 1 > gpurun_out/n${N}_bench_ref.json 2> gpurun_out/n${N}_bench_ref.err; echo "ref arm rc=$?"; tail -1 gpurun_out/n${N}_bench_ref.json | cut -c1-260
