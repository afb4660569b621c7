timeout 900 python -m pytest tests/test_step_gpu.py -q --timeout 200 -k conditional 2>&1 | grep -E "^E   |passed|failed" | cut -c1-300 | head -20
