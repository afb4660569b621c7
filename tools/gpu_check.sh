mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x 2>&1 | grep -E "^E   .*(Assert|Error)|passed|failed|^FAILED|Error" | cut -c1-250 | head -20
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline 2>gpurun_out/bq_err.log | tee gpurun_out/bench_quick.json | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print(d['e2e']['value'], d['e2e']['last_loss']); print({k:round(v,4) for k,v in d['kernels_ms'].items()}); print(d['roofline']['frac'], d['clocks'])"
tail -3 gpurun_out/bq_err.log | cut -c1-300
