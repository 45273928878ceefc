mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_edges.py -q 2>&1 | tail -15
timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:chol_factor_kernel -c 1 -o gpurun_out/chol_factor python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_chol.log 2>&1; tail -1 gpurun_out/ncu_chol.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:kmat_rows_kernel -s 6 -c 1 -o gpurun_out/kmat_rows python bench.py --workload kf_init --steps 2 --warmup 1 > gpurun_out/ncu_kmat.log 2>&1; tail -1 gpurun_out/ncu_kmat.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_kfinit.csv python scripts/kfinit_once.py > gpurun_out/kfinit_once.log 2>&1; tail -2 gpurun_out/kfinit_once.log
