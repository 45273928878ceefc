#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_solve.py -q -x 2>&1 | tail -3
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"chol_factor2" -c 1 -f -o gpurun_out/chol_r02_factor2 \
  python bench.py --workload ba_window --steps 1 --warmup 3 --no-e2e 1 > gpurun_out/aj_ncu.log 2>&1
python scripts/ncu_summary.py gpurun_out/chol_r02_factor2.ncu-rep gpurun_out/aj_chol_factor2_full.txt
cat gpurun_out/aj_chol_factor2_full.txt
