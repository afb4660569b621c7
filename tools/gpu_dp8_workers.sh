mkdir -p gpurun_out
nproc > gpurun_out/nproc8.txt; lscpu | grep -i "model name\|socket\|numa node(s)\|^CPU(s)" >> gpurun_out/nproc8.txt
for W in 3 4; do
BENCH_WORKERS=$W BENCH_WATCHDOG=170 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2971$W bench.py --gpus 8 --steps 60 --warmup 5 --no-cpu-baseline --also-config3 0 2>gpurun_out/dp_err_8_w$W.log | tee gpurun_out/scale_r2_8_w$W.json | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N', d['n_gpus'], {k:round(d[k],3) for k in ('value','ms_per_step')}, 'e2e', round(d['e2e']['value']), d['e2e']['host_side'][:60])"
done
cat gpurun_out/nproc8.txt
