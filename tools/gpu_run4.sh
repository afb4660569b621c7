mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_step_gpu.py -q --timeout 120 2>&1 | tail -30 > gpurun_out/t.log
python tools/summarize_fail.py gpurun_out/t.log
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>/dev/null | tee gpurun_out/bench_quick.json | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print(d['e2e']); print(d['kernels_ms']); print(d['roofline']['frac'], d['spmm'])"
