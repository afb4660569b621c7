mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/t_all.log 2>&1
grep -E "passed|failed|^E  |^FAILED" gpurun_out/t_all.log | cut -c1-300 | head -30
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log | cut -c1-200
BENCH_WATCHDOG=400 timeout 500 python bench.py 2>gpurun_out/bench1_err.log | tee gpurun_out/bench_r2_1.json | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N', d['n_gpus'], {k:round(d[k],3) for k in ('value','ms_per_step')}, 'e2e', {k:v for k,v in d['e2e'].items() if k!='host_side'}); print(d['roofline']); print(d.get('cpu_baseline')); print(d.get('parity_check',{}).get('rel_err')); print(d['gpu_launches'], d['clocks'])"
grep -E "Error|Traceback" -A8 gpurun_out/bench1_err.log | head -30
timeout 400 python bench.py --impl reference --steps 4 --warmup 1 2>/dev/null | tee gpurun_out/bench_r2_ref.json | cut -c1-600
