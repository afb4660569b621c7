mkdir -p gpurun_out
run() { # n extra-env tag flags
n=$1; tag=$3
env $2 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $n --steps 40 --warmup 5 --no-cpu-baseline $4 2>gpurun_out/dp_err_$tag.log | tee gpurun_out/scale_$tag.json | python -c "
import json,sys,os; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$tag N', d['n_gpus'], {k:round(d[k],3) for k in ('value','ms_per_step')}, round(d['e2e']['value']), {k:round(v,3) for k,v in d['kernels_ms'].items()})"
grep -E "Error|Traceback" -A3 gpurun_out/dp_err_$tag.log | head -10
}
run 4 "NCCL_MAX_CTAS=16 CMMVAE_NCCL_SMS=16" 4c16 ""
run 4 "NCCL_MAX_CTAS=8 CMMVAE_NCCL_SMS=8" 4c8 ""
