# usage: gpu_dpN.sh N [steps]
N=${1:-2}; STEPS=${2:-60}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
if [ "$N" = "2" ] && [ -z "$SKIP_TEST" ]; then
timeout 600 python -X faulthandler -u -m pytest tests/test_dp_gpu.py -m gpu -v --timeout 500 -k "two_gpus" > gpurun_out/t_dp2gpu.log 2>&1
grep -E "PASSED|FAILED|SKIPPED|passed|failed|^E  " gpurun_out/t_dp2gpu.log | cut -c1-300 | head -20
fi
BENCH_PINNED_LEG=${PINNED:-0} BENCH_WATCHDOG=170 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $N --steps $STEPS --warmup 5 --no-cpu-baseline 2>gpurun_out/dp_err_$N.log | tee gpurun_out/scale_r2_$N.json | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N', d['n_gpus'], {k:round(d[k],3) for k in ('value','ms_per_step')}, 'e2e', round(d['e2e']['value']), d['e2e'].get('host_ms_per_step'), 'pinned', (d['e2e'].get('pinned_chunks') or {}).get('value')); print({k:round(v,3) for k,v in d['kernels_ms'].items()}); print(d.get('parity_check',{}).get('rel_err')); print(d.get('also'))"
grep -E "Error|Traceback|error" -A5 gpurun_out/dp_err_$N.log | grep -v "NCCL INFO" | head -30
