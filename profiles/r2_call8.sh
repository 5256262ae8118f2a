mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py -m gpu -q -x > gpurun_out/r2_gemm_tests.log 2>&1
tail -30 gpurun_out/r2_gemm_tests.log
timeout 900 python -m pytest tests/test_gpu_pdf.py tests/test_gpu_summarizers.py -m gpu -q > gpurun_out/r2_misc_tests.log 2>&1
tail -8 gpurun_out/r2_misc_tests.log
