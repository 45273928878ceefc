#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for B in 74 1; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:track_pyr -s 3 -c 1 -f -o gpurun_out/trk_b$B \
  python bench.py --workload track640 --batch $B --steps 1 --warmup 3 --no-e2e 1 > gpurun_out/b_ncu_b$B.log 2>&1
tail -3 gpurun_out/b_ncu_b$B.log
done
ls -la gpurun_out/*.ncu-rep
