mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_step_gpu.py -q --timeout 300 2>&1 | tail -400 > gpurun_out/step_test.log
tail -50 gpurun_out/step_test.log
