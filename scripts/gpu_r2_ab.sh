#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python bench.py --workload sfm --steps 5 --warmup 2 > gpurun_out/ab_sfm.json 2>gpurun_out/ab_sfm.err
echo "sfm rc=$?"; tail -c 1500 gpurun_out/ab_sfm.json; tail -c 800 gpurun_out/ab_sfm.err
