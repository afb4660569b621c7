python tools/time_dec.py
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_step_gpu.py -q --timeout 120 2>&1 | tail -3
