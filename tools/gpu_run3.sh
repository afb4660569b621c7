mkdir -p gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -5
python bench.py --steps 20 --warmup 5 2>gpurun_out/bench_err.log | tee gpurun_out/bench_r1_first.json
tail -5 gpurun_out/bench_err.log
