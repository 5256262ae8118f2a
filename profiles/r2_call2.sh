mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_persistent.py -m gpu -x -q > gpurun_out/r2_persist_tests.log 2>&1
tail -30 gpurun_out/r2_persist_tests.log
timeout 600 python -m pytest tests/test_gpu_mdn.py tests/test_gpu_bayessim.py -m gpu -q > gpurun_out/r2_mdn_tests.log 2>&1
tail -8 gpurun_out/r2_mdn_tests.log
timeout 300 python profiles/time_breakdown.py > gpurun_out/r2_breakdown_persistent.txt 2>&1
cat gpurun_out/r2_breakdown_persistent.txt | tail -14
timeout 600 compute-sanitizer --tool memcheck --log-file gpurun_out/r2_persist_memcheck.san python -m pytest tests/test_gpu_persistent.py -m gpu -q -k "oracle and (31 or MDRFF-23)" > gpurun_out/r2_persist_memcheck.log 2>&1
tail -3 gpurun_out/r2_persist_memcheck.log; tail -5 gpurun_out/r2_persist_memcheck.san
