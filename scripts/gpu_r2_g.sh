#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_track.py tests/test_gpu_tracking_class.py -q -m gpu > gpurun_out/g_track_tests.log 2>&1
echo "track tests rc=$?" >> gpurun_out/g_track_tests.log
grep -n "^E  .*it [0-9]\|passed\|failed\|out of bounds" gpurun_out/g_track_tests.log | cut -c1-330 | tail -60
