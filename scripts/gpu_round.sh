mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_ba_n1.json 2> gpurun_out/bench_ba_n1.err; cut -c1-400 gpurun_out/bench_ba_n1.json
timeout 600 python bench.py --workload track640 > gpurun_out/bench_track_n1.json 2> gpurun_out/bench_track_n1.err; cut -c1-300 gpurun_out/bench_track_n1.json
timeout 600 python bench.py --workload kf_init --steps 10 --warmup 3 > gpurun_out/bench_kfinit_n1.json 2> gpurun_out/bench_kfinit_n1.err; cut -c1-300 gpurun_out/bench_kfinit_n1.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_ba.json 2> gpurun_out/bench_ref_ba.err; cut -c1-300 gpurun_out/bench_ref_ba.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/launches_ba_final.csv python bench.py --steps 2 --warmup 1 > /dev/null 2>&1
