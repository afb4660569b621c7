timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_step_gpu.py tests/test_modules_gpu.py -q --timeout 200 2>&1 | grep -E "^E   |passed|failed" | cut -c1-200 | head -20
