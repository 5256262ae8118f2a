mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_corr_layer.py -m gpu -q -x > gpurun_out/r2c_gemm_tests.log 2>&1
tail -3 gpurun_out/r2c_gemm_tests.log
timeout 600 python profiles/rooflines_only.py 2>&1 | grep "rff\|corr_fused"
timeout 300 python profiles/corr_dbg.py 2>&1 | tail -3
