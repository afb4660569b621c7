# ncu evidence for profiles/: (1) launch list with per-launch device time, (2) full capture of the top kernels
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench1.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k regex:"decoder_mse_fused_kernel|spmm_tc_kernel|gemm_bf16_tc_kernel|clip_adam_kernel|csr_linear_fwd_kernel" \
    -s 14 -c 7 -f -o gpurun_out/topk python tools/prof_kernels.py > gpurun_out/ncu_topk.log 2>&1
tail -2 gpurun_out/ncu_topk.log
