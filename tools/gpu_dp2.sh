mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dp_gpu.py tests/test_step_gpu.py -q --timeout 600 -x 2>&1 | grep -E "^E   |passed|failed|^FAILED|Error" | cut -c1-300 | head -30
for m in 1 0; do
CMMVAE_DP_BY_INPUTS=$m timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 40 --warmup 5 --no-cpu-baseline 2>gpurun_out/dp_err_2_$m.log | tee gpurun_out/scale_2_$m.json | python -c "
import json,sys,os; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('by_inputs=$m N', d['n_gpus'], {k:round(d[k],3) for k in ('value','ms_per_step')}, round(d['e2e']['value']), d['e2e']['last_loss'], {k:round(v,3) for k,v in d['kernels_ms'].items()})"
grep -E "Error|error|Traceback" -A3 gpurun_out/dp_err_2_$m.log | head -20
done
