#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 600 python -m pytest tests/test_gpu_ba.py -q -m gpu -k "three_exchange or golden" > gpurun_out/i_ba_tests.log 2>&1
tail -3 gpurun_out/i_ba_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_ba_check.py > gpurun_out/i_dist_check.log 2>&1
grep -n "iter\|DIST_BA_CHECK\|Error" gpurun_out/i_dist_check.log | tail -8
COMO_B200_SHARD_MEDIAN=three timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/dist_ba_check.py > gpurun_out/i_dist_check_legacy.log 2>&1
grep -n "DIST_BA_CHECK\|Error" gpurun_out/i_dist_check_legacy.log | tail -3
for mode in six three; do
COMO_B200_SHARD_MEDIAN=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 --workload ba_window --shard 1 --no-e2e 1 > gpurun_out/i_shard2_$mode.json 2> gpurun_out/i_shard2_$mode.err
tail -c 900 gpurun_out/i_shard2_$mode.json; tail -3 gpurun_out/i_shard2_$mode.err
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/i_bench_all_n2.json 2> gpurun_out/i_bench_all_n2.err
tail -c 600 gpurun_out/i_bench_all_n2.json; tail -3 gpurun_out/i_bench_all_n2.err
