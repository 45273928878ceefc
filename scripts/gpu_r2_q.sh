#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python scripts/stream_ctas_check.py 2>&1 | tail -8
for c in 0 -1; do
COMO_B200_STREAM_CTAS=$c timeout 600 python -m pytest tests/test_gpu_ba.py -q -m gpu -k "k32 or k8_window" 2>&1 | tail -3
done
