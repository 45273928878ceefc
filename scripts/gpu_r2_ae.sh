#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/ae_pytest_gpu.log 2>&1
echo "pytest -m gpu rc=$?" >> gpurun_out/ae_pytest_gpu.log
tail -12 gpurun_out/ae_pytest_gpu.log
