mkdir -p gpurun_out
run() { # n extra-env tag flags
n=$1; tag=$3
env $2 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $n --steps 40 --warmup 5 --no-cpu-baseline $4 2>gpurun_out/dp_err_$tag.log | tee gpurun_out/scale_$tag.json | python -c "
import json,sys,os; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$tag N', d['n_gpus'], {k:round(d[k],3) for k in ('value','ms_per_step')}, round(d['e2e']['value']), d.get('e2e_diag'), {k:round(v,3) for k,v in d['kernels_ms'].items() if k.startswith('dp_') or k.startswith('norm')})"
grep -m3 -i "nvls\|Using network\|Connected all" gpurun_out/dp_err_$tag.log | cut -c1-200
}
run 4 "NCCL_NVLS_ENABLE=0" 4nonvls "--e2e-diag"
run 4 "BENCH_NCCL_DEBUG=INFO" 4info "--e2e-diag"
