# ncu evidence for profiles/ (round 2, final state): launch list over the bench command + full captures of the top kernels
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-torch-baseline --no-parity-check > gpurun_out/ncu_bench1.log 2>&1
tail -1 gpurun_out/ncu_bench1.log | cut -c1-200
ncu --set full --clock-control none --import-source on \
    -k regex:"decoder_mse_fused_kernel|spmm_tc_kernel|gemm_bf16_tc_kernel|clip_adam_kernel|csr_linear_fwd_kernel" \
    -s 14 -c 7 -f -o gpurun_out/topk python tools/prof_kernels.py > gpurun_out/ncu_topk.log 2>&1
tail -1 gpurun_out/ncu_topk.log
ls -la gpurun_out/*.ncu-rep
