#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_track.py tests/test_gpu_tracking_class.py tests/test_gpu_edges.py -x -q > gpurun_out/w_tests.log 2>&1
echo "track tests rc=$?"; tail -15 gpurun_out/w_tests.log
run() {  # label, batch, env...
  label=$1; batch=$2; shift; shift
  env "$@" timeout 300 python bench.py --batch $batch --workload track640 --steps 10 --warmup 3 --no-e2e 1 > gpurun_out/w_$label.json 2>gpurun_out/w_$label.err
  python - "$label" <<'PY'
import json,sys
try:
    d=json.loads(open('gpurun_out/w_%s.json'%sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d['value']), 'it/s frac', round(d['roofline']['frac'],4), 'ms', round(d['roofline']['launch_ms'],4), )
except Exception as e:
    print(sys.argv[1], 'FAILED', e); print(open('gpurun_out/w_%s.err'%sys.argv[1]).read()[-600:])
PY
}
run b592 592 A=1
run b444 444 A=1
run b296 296 A=1
