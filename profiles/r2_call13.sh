mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_corr_layer.py -m gpu -q -x > gpurun_out/r2b_corr_tests.log 2>&1
tail -30 gpurun_out/r2b_corr_tests.log
