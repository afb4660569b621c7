# ncu evidence for profiles/ (round 2): (1) launch list with per-launch device time over the bench command,
# (2) full captures of the top kernels at the bench shape, (3) the fused decoder at the config 3 / 4 batch sizes
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-torch-baseline --no-parity-check > gpurun_out/ncu_bench1.log 2>&1
tail -1 gpurun_out/ncu_bench1.log | cut -c1-200
ncu --set full --clock-control none --import-source on \
    -k regex:"decoder_mse_fused_kernel|spmm_tc_kernel|gemm_bf16_tc_kernel|clip_adam_kernel|csr_linear_fwd_kernel" \
    -s 14 -c 7 -f -o gpurun_out/topk python tools/prof_kernels.py > gpurun_out/ncu_topk.log 2>&1
tail -1 gpurun_out/ncu_topk.log
for PB in 4096 8192; do
PB=$PB ncu --set full --clock-control none -k regex:"decoder_mse_fused_kernel" -s 2 -c 1 -f -o gpurun_out/dec$PB \
    python tools/prof_kernels.py > gpurun_out/ncu_dec$PB.log 2>&1
tail -1 gpurun_out/ncu_dec$PB.log
done
