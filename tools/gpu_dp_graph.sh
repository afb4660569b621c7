mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_dp_gpu.py tests/test_graph_gpu.py -m gpu -q --timeout 400 -x 2>&1 | tail -25 | cut -c1-220
BENCH_SAME_GPU=1 BENCH_WATCHDOG=150 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29713 bench.py --gpus 2 --steps 30 --warmup 4 --no-cpu-baseline --also-config3 0 2>gpurun_out/dp_same_err.log | tee gpurun_out/dp_same.json | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N', d['n_gpus'], {k:round(d[k],3) for k in ('value','ms_per_step')}, 'e2e', round(d['e2e']['value']), d['e2e']['last_loss']); print({k:round(v,3) for k,v in d['kernels_ms'].items()}); print(d.get('parity_check',{}).get('rel_err'))"
grep -v "NCCL INFO\|Warning\|warn\|sparse_csr\|^\*\*\*\|OMP_NUM" gpurun_out/dp_same_err.log | head -40 | cut -c1-200
