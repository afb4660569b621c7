mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q --timeout 300 -k "tf32" > gpurun_out/t_tf32.log 2>&1
grep -E "passed|failed|^E  |^FAILED" gpurun_out/t_tf32.log | cut -c1-300 | head -12
timeout 900 python -X faulthandler -u -m pytest tests/test_dp_gpu.py -m gpu -v --timeout 600 -x > gpurun_out/t_dp.log 2>&1
echo "rc=$?" >> gpurun_out/t_dp.log
grep -E "PASSED|FAILED|passed|failed|rc=|^E  " gpurun_out/t_dp.log | cut -c1-400 | head -30
timeout 600 python -m pytest tests/test_step_gpu.py tests/test_engine_state_gpu.py tests/test_modules_gpu.py -m gpu -q --timeout 600 > gpurun_out/t_step.log 2>&1
grep -E "passed|failed|^E  |^FAILED" gpurun_out/t_step.log | cut -c1-300 | head -20
python tools/diag_precision.py base > gpurun_out/diag_precision2.log 2>&1; grep -A16 "== base" gpurun_out/diag_precision2.log
