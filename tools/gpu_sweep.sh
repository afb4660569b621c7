mkdir -p gpurun_out
timeout 600 python tools/spmm_sweep.py > gpurun_out/spmm_sweep.log 2>&1; tail -2 gpurun_out/spmm_sweep.log | cut -c1-200
timeout 300 python tools/spmm_sweep.py --shard 8 --batches 512 1024 4096 > gpurun_out/spmm_sweep8.log 2>&1; tail -1 gpurun_out/spmm_sweep8.log | cut -c1-200
timeout 200 python tools/spmm_sweep.py --shard 2 --batches 1024 4096 --densities 0.01 0.05 0.2 > gpurun_out/spmm_sweep2.log 2>&1; tail -1 gpurun_out/spmm_sweep2.log | cut -c1-200
