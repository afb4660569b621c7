mkdir -p gpurun_out
(lscpu | grep -i "model name\|socket\|core\|thread\|numa\|L3\|^CPU(s)"; free -g | head -2; ulimit -l; cat /sys/kernel/mm/transparent_hugepage/enabled) > gpurun_out/hostinfo.txt 2>&1
cat gpurun_out/hostinfo.txt
PYTHONPATH=. python tools/host_pack_scaling.py 1 4 8 16 24 32 2>&1 | tee gpurun_out/host_pack_scaling.txt
