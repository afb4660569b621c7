mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py -q --timeout 120 2>&1 | tail -60 > gpurun_out/ops_test.log
tail -45 gpurun_out/ops_test.log
