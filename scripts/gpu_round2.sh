mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_edges.py -q 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_ba_n2.json 2> gpurun_out/bench_ba_n2.err; cut -c1-420 gpurun_out/bench_ba_n2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --shard 1 --steps 20 --warmup 5 > gpurun_out/bench_ba_n2_shard.json 2> gpurun_out/bench_ba_n2_shard.err; cut -c1-420 gpurun_out/bench_ba_n2_shard.json; tail -3 gpurun_out/bench_ba_n2_shard.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 scripts/dist_ba_check.py > gpurun_out/dist_check.log 2>&1; tail -4 gpurun_out/dist_check.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --workload track640 > gpurun_out/bench_track_n2.json 2> gpurun_out/bench_track_n2.err; cut -c1-300 gpurun_out/bench_track_n2.json
