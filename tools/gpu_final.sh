mkdir -p gpurun_out
timeout 1200 python tools/spmm_sweep.py > gpurun_out/spmm_sweep.log 2>&1; tail -2 gpurun_out/spmm_sweep.log
bash tools/gpu_profile.sh
timeout 600 python bench.py --steps 200 --warmup 10 2>gpurun_out/bench_err.log | tee gpurun_out/bench_r1.json | cut -c1-400
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>>gpurun_out/bench_err.log | tee gpurun_out/bench_r1_reference.json | cut -c1-300
