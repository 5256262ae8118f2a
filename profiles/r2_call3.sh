mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_persistent.py -m gpu -q > gpurun_out/r2_persist_tests.log 2>&1
tail -25 gpurun_out/r2_persist_tests.log
timeout 600 python -m pytest tests/test_gpu_mdn.py tests/test_gpu_bayessim.py -m gpu -q > gpurun_out/r2_mdn_tests.log 2>&1
tail -5 gpurun_out/r2_mdn_tests.log
timeout 300 python profiles/tp_profile.py > gpurun_out/r2_tp_profile.txt 2>&1
cat gpurun_out/r2_tp_profile.txt | tail -26
