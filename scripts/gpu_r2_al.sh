#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 600 python -m pytest tests/test_gpu_ba.py tests/test_gpu_edges.py -q -x 2>&1 | grep -v "^$" | grep -v "^tensor\|^        \[" | tail -45
