mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_corr_layer.py -m gpu -q -x -k "cartpole or halfcheetah or ant-100 or envelope" > gpurun_out/r2c_corr_memcheck.log 2>&1
tail -6 gpurun_out/r2c_corr_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_corr_layer.py -m gpu -q -x -k "weight_gradient and cartpole or forward and cartpole" > gpurun_out/r2c_corr_racecheck.log 2>&1
tail -6 gpurun_out/r2c_corr_racecheck.log
