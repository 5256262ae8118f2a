mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py -m gpu -q -x > gpurun_out/r2c_gemm_tests.log 2>&1
tail -4 gpurun_out/r2c_gemm_tests.log
timeout 600 python profiles/rooflines_only.py 2>&1 | grep "rff"
BSIG_TC_M2=0 timeout 600 python profiles/rooflines_only.py 2>&1 | grep "rff_projection_ant_64k_tcgen05_tf32 "
