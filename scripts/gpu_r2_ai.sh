#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 300 python scripts/chol_timeline2.py 2>&1 | tail -10
timeout 300 python -m pytest tests/test_gpu_solve.py -q -x 2>&1 | tail -2
