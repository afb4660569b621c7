mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/t_all.log 2>&1
grep -E "passed|failed|^E  |^FAILED" gpurun_out/t_all.log | cut -c1-300 | head -30
