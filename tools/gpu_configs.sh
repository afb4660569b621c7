mkdir -p gpurun_out
for c in 3 4; do
timeout 300 python bench.py --config $c --steps 30 --warmup 4 --no-cpu-baseline 2>gpurun_out/c${c}_err.log | tee gpurun_out/bench_config$c.json | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['config']['workload']); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, round(d['e2e']['value']), d['roofline']['frac'])"
tail -2 gpurun_out/c${c}_err.log | cut -c1-200
done
