mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_step_gpu.py -m gpu -q --timeout 500 -k "conditional" > gpurun_out/t_cond.log 2>&1
grep -E "passed|failed|^E  |^FAILED|Error" gpurun_out/t_cond.log | cut -c1-300 | head -40
