mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kfinit.py tests/test_gpu_depthcov.py -q 2>&1 | tail -40
timeout 300 python scripts/time_kmat.py 2>&1 | tail -12
