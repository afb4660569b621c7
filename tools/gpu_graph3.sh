mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_graph_gpu.py tests/test_step_gpu.py -m gpu -q --timeout 300 2>&1 | grep -E "passed|failed|^E  " | cut -c1-200 | head -6
BENCH_WATCHDOG=200 timeout 280 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-torch-baseline 2>gpurun_out/bench_err.log | tee gpurun_out/bench_r2_d.json | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print(d['e2e']['value'], d['e2e']['last_loss']); print({k:round(v,4) for k,v in d['kernels_ms'].items()}); print(d['roofline']['frac'], d.get('parity_check',{}).get('rel_err'))"
grep -v "Warn\|warn" gpurun_out/bench_err.log | tail -5 | cut -c1-300
