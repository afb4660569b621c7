mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"spmm_tc_kernel|decoder_mse_fused_kernel|gemm_bf16_tc_kernel" \
    -s 10 -c 5 -f -o gpurun_out/prof2 python tools/prof_kernels.py > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/ncu2.log
ls -la gpurun_out/prof2.ncu-rep
