#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_track.py tests/test_gpu_tracking_class.py tests/test_gpu_edges.py -q > gpurun_out/y_tests.log 2>&1
echo "track tests rc=$?"; tail -4 gpurun_out/y_tests.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:track_pyr -s 3 -c 1 -f -o gpurun_out/trk_r02_b592 \
  python bench.py --workload track640 --batch 592 --steps 1 --warmup 3 --no-e2e 1 > gpurun_out/y_ncu.log 2>&1
tail -2 gpurun_out/y_ncu.log | cut -c1-300
ls -la gpurun_out/trk_r02_b592.ncu-rep
python scripts/ncu_summary.py gpurun_out/trk_r02_b592.ncu-rep gpurun_out/y_trk_b592_full.txt
ncu -i gpurun_out/trk_r02_b592.ncu-rep --page raw --csv > gpurun_out/y_trk_b592_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/y_trk_b592_raw.csv')))
h,u,r=rows[0],rows[1],rows[2]
for k in ["gpu__time_duration.sum","dram__bytes_read.sum","dram__bytes_write.sum","dram__throughput.avg.pct_of_peak_sustained_elapsed","l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct","lts__t_sector_hit_rate.pct","sm__inst_executed.sum","smsp__inst_executed.avg.per_cycle_active","sm__warps_active.avg.pct_of_peak_sustained_active","launch__registers_per_thread","launch__occupancy_limit_registers","smsp__thread_inst_executed.sum","sm__throughput.avg.pct_of_peak_sustained_elapsed"]:
    if k in h: print(k, r[h.index(k)], u[h.index(k)])
PY
timeout 300 python bench.py --workload track640 --batch 1 --steps 20 --warmup 3 > gpurun_out/y_b1_e2e.json 2>gpurun_out/y_b1_e2e.err
python -c "
import json
d=json.loads(open('gpurun_out/y_b1_e2e.json').read().strip().splitlines()[-1]); print('B=1 e2e', d['e2e'], 'kernel ms', d['roofline']['launch_ms'])"
