mkdir -p gpurun_out
for ch in 4 8 16; do
export NCCL_MAX_NCHANNELS=$ch
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu-baseline 2>gpurun_out/dp_err_2.log | python -c "
import json,sys,os; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('nch', os.environ['NCCL_MAX_NCHANNELS'], d['n_gpus'], {k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['kernels_ms'])"
done
