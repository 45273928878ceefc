mkdir -p gpurun_out
for b in 37 74; do
timeout 400 python bench.py --workload track640 --batch $b --steps 10 --warmup 3 > gpurun_out/bench_track_b$b.json 2> gpurun_out/bench_track_b$b.err; python -c "import json; d=json.loads(open('gpurun_out/bench_track_b$b.json').read()); print($b, d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'])" || tail -3 gpurun_out/bench_track_b$b.err
done
