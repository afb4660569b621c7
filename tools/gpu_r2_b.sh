mkdir -p gpurun_out
timeout 1200 python -X faulthandler -u -m pytest tests/test_parity_fullshape_gpu.py tests/test_engine_state_gpu.py -m gpu -v --timeout 900 > gpurun_out/t_new.log 2>&1
echo "rc=$?" >> gpurun_out/t_new.log
grep -E "passed|failed|rc=" gpurun_out/t_new.log | tail -3
timeout 900 python -m pytest tests -m gpu -q --timeout 600 --deselect tests/test_parity_fullshape_gpu.py > gpurun_out/t_all.log 2>&1
grep -E "passed|failed" gpurun_out/t_all.log | tail -2
dmesg 2>/dev/null | tail -5 > gpurun_out/dmesg.txt
