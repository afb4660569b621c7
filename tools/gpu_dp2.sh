mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 400 2>&1 | grep -E "^E   |passed|failed|skipped" | cut -c1-250 | head
for n in 1 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $n --steps 50 --warmup 5 --no-cpu-baseline 2>gpurun_out/dp_err_$n.log | tee gpurun_out/scale_$n.json | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['n_gpus'], {k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['kernels_ms'])"
done
