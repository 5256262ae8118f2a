mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_corr_layer.py -m gpu -q -x -k "weight_gradient" > gpurun_out/r2b_corr_wg.log 2>&1
tail -12 gpurun_out/r2b_corr_wg.log
timeout 300 python -m pytest tests/test_gpu_corr_layer.py -m gpu -q -x -k "adam_epilogue" > gpurun_out/r2b_corr_adam.log 2>&1
tail -12 gpurun_out/r2b_corr_adam.log
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_corr_layer.py -m gpu -q -x -k "fused_forward and cartpole-100" > gpurun_out/r2b_corr_fwd_san.log 2>&1
grep -v "^$" gpurun_out/r2b_corr_fwd_san.log | grep -A12 "=========" | head -60
