mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_graph_gpu.py tests/test_dp_gpu.py tests/test_step_gpu.py tests/test_output_discriminator_gpu.py tests/test_engine_state_gpu.py -m gpu -q --timeout 500 2>&1 | tail -4 | cut -c1-200
BENCH_WATCHDOG=300 timeout 400 python bench.py --no-cpu-baseline --no-torch-baseline 2>gpurun_out/b_err.log | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; print(round(d['value']), 'e2e', round(e['value']), e['host_ms_per_step'], d['clocks'], d.get('parity_check',{}).get('rel_err'))"
grep -E "Error" -A5 gpurun_out/b_err.log | head -8
