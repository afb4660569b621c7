mkdir -p gpurun_out
run() { # n tag
n=$1; tag=$2
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $n --steps 60 --warmup 5 --no-cpu-baseline 2>gpurun_out/dp_err_$tag.log | tee gpurun_out/scale_$tag.json | python -c "
import json,sys,os; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N', d['n_gpus'], {k:round(d[k],3) for k in ('value','ms_per_step')}, round(d['e2e']['value']), {k:round(v,3) for k,v in d['kernels_ms'].items()})"
grep -E "Error|Traceback" -A3 gpurun_out/dp_err_$tag.log | head -10
}
for n in ${SCALE_NS:-8 4 2}; do run $n $n; done
