mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/t_all.log 2>&1
grep -E "passed|failed|^E  |^FAILED|ELBO curve" gpurun_out/t_all.log | cut -c1-300 | head -30
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-torch-baseline 2>gpurun_out/bench_err.log | tee gpurun_out/bench_r2_b.json | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print(d['e2e']['value'], d['e2e']['last_loss']); print({k:round(v,4) for k,v in d['kernels_ms'].items()}); print(d['roofline']['frac'], d.get('parity_check',{}).get('rel_err'))"
tail -3 gpurun_out/bench_err.log | cut -c1-300
