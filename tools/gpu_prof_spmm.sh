mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"spmm_tc_kernel" -s 2 -c 2 -f -o gpurun_out/spmm python tools/prof_kernels.py > gpurun_out/ncu_spmm.log 2>&1
tail -2 gpurun_out/ncu_spmm.log
