mkdir -p gpurun_out
for n in 8 4; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $n --steps 50 --warmup 5 --no-cpu-baseline 2>gpurun_out/dp_err_$n.log | tee gpurun_out/scale_$n.json | python -c "
import json,sys,os; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N', d['n_gpus'], {k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['kernels_ms'])"
done
