mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 300 2>&1 | tail -40 > gpurun_out/gpu_tests.log
python tools/summarize_fail.py gpurun_out/gpu_tests.log
python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>/dev/null | tee gpurun_out/bench_quick.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print(d['e2e']); print(d['kernels_ms']); print(d['roofline']['frac'], d['spmm'])"
