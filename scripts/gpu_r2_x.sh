#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python scripts/track_batch_debug.py 444 2>&1 | tail -5
timeout 900 compute-sanitizer --tool memcheck --print-limit 6 python scripts/track_batch_debug.py 230 > gpurun_out/x_san.log 2>&1
echo "sanitizer rc=$?"; grep -v "^$" gpurun_out/x_san.log | grep -v "Host Frame" | head -60
