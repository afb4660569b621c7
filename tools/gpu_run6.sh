mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -q --timeout 120 -k decoder 2>&1 | tail -2
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>/dev/null | tee gpurun_out/bench_quick.json | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print(d['e2e']['value']); print(d['kernels_ms']); print(d['roofline']['frac'])"
ncu --metrics gpu__time_duration.sum --clock-control none -s 260 -c 130 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_bench1.log 2>&1
