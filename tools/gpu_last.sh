mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_step_gpu.py tests/test_modules_gpu.py tests/test_reference_unit_tests_gpu.py -x -q -m gpu --timeout 300 2>&1 | grep -E "passed|failed|^E  |^FAILED" | cut -c1-200 | head
BENCH_WATCHDOG=400 timeout 500 python bench.py 2>gpurun_out/bench1_err.log | tee gpurun_out/bench_r2_1.json | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N', d['n_gpus'], {k:round(d[k],3) for k in ('value','ms_per_step')}, 'e2e', round(d['e2e']['value']), d['roofline']['frac'], d['clocks'], d.get('parity_check',{}).get('rel_err'))"
grep -E "Error|Traceback" -A8 gpurun_out/bench1_err.log | head -20
