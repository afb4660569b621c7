mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_step_gpu.py tests/test_feed.py -q --timeout 300 2>&1 | grep -E "^E   |passed|failed|^FAILED" | cut -c1-250 | head -20
for c in 2 4; do
timeout 900 python bench.py --config $c --steps 30 --warmup 4 --no-cpu-baseline 2>gpurun_out/c${c}_err.log | tee gpurun_out/bench_config$c.json | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['config']['workload']); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print(d['e2e']); print(d['kernels_ms']); print(d['roofline']['frac'])"
tail -3 gpurun_out/c${c}_err.log | cut -c1-300
done
