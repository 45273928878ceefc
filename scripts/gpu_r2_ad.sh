#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"ba_residual|ba_accum" -c 2 -f -o gpurun_out/ba_r02_photo \
  python bench.py --workload ba_window --steps 1 --warmup 3 --no-e2e 1 > gpurun_out/ad_ncu.log 2>&1
tail -2 gpurun_out/ad_ncu.log | cut -c1-200
ls -la gpurun_out/ba_r02_photo.ncu-rep
python scripts/ncu_summary.py gpurun_out/ba_r02_photo.ncu-rep gpurun_out/ad_ba_photo_full.txt
cat gpurun_out/ad_ba_photo_full.txt
ncu -i gpurun_out/ba_r02_photo.ncu-rep --page raw --csv > gpurun_out/ad_ba_photo_raw.csv 2>/dev/null
