mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_corr_layer.py -m gpu -q -x > gpurun_out/r2b_corr_tests.log 2>&1
tail -3 gpurun_out/r2b_corr_tests.log
timeout 300 python profiles/corr_dbg.py 2>&1 | tail -8
