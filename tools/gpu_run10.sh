mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_step_gpu.py tests/test_modules_gpu.py tests/test_feed.py -q --timeout 300 2>&1 | grep -E "^E   .*(Assert|Error)|passed|failed|^FAILED" | cut -c1-250 | head -20
for pdl in 0 1; do
CMMVAE_PDL=$pdl timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline 2>gpurun_out/bq_err.log | tee gpurun_out/bench_pdl$pdl.json | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('PDL=$pdl', {k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print(d['e2e']['value']); print({k:round(v,4) for k,v in d['kernels_ms'].items()})"
tail -3 gpurun_out/bq_err.log | cut -c1-300
done
