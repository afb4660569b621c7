mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/env2.txt; nproc >> gpurun_out/env2.txt; free -g | head -2 >> gpurun_out/env2.txt
timeout 900 python -m pytest tests/test_parity_fullshape_gpu.py tests/test_engine_state_gpu.py tests/test_feed.py -m gpu -q --timeout 600 2>&1 | tail -60 > gpurun_out/t_new.log
grep -E "passed|failed" gpurun_out/t_new.log | tail -2
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x --deselect tests/test_parity_fullshape_gpu.py 2>&1 | tail -30 > gpurun_out/t_all.log
grep -E "passed|failed" gpurun_out/t_all.log | tail -2
timeout 600 python bench.py --steps 100 --warmup 5 2>gpurun_out/bench_err.log | tee gpurun_out/bench_r2_a.json | cut -c1-600
tail -5 gpurun_out/bench_err.log | cut -c1-300
timeout 600 python tools/library_baselines.py > gpurun_out/library_baselines.log 2>&1; tail -3 gpurun_out/library_baselines.log | cut -c1-400
