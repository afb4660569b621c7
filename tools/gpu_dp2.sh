mkdir -p gpurun_out
run() {
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu-baseline 2>gpurun_out/dp_err_2.log | python -c "
import json,sys,os; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernels_ms']; print(os.environ.get('TAG'), round(d['ms_per_step'],3), {x: round(k[x],3) for x in ('dh_gemm','csr_linear_fwd','norm+clip_adam','dp_wait_shadow_first','dp_wait_grads')})"
}
TAG=default run
TAG=min32 NCCL_MIN_NCHANNELS=32 run
TAG=min32_sms40 NCCL_MIN_NCHANNELS=32 CMMVAE_NCCL_SMS=40 run
TAG=sms8 CMMVAE_NCCL_SMS=8 run
TAG=max8_sms8 NCCL_MAX_NCHANNELS=8 CMMVAE_NCCL_SMS=8 run
