mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_graph_gpu.py -m gpu -q --timeout 300 2>&1 | grep -E "passed|failed|^E  " | cut -c1-200 | head -6
for mode in "BENCH_GRAPH=1" "BENCH_GRAPH=0" "BENCH_GRAPH=1 CMMVAE_PDL=0" "BENCH_GRAPH=0 CMMVAE_PDL=0"; do
env $mode BENCH_WATCHDOG=100 timeout 150 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-torch-baseline --no-parity-check 2>gpurun_out/bench_err.log | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$mode', {k:round(d[k],4) for k in ('value','ms_per_step')}, 'e2e', round(d['e2e']['value']), 'dec', round(d['roofline']['ms_per_launch'],4), 'spmm', round(d['spmm']['ms'],4))"
done
grep -v "Warn\|warn" gpurun_out/bench_err.log | tail -5 | cut -c1-300
