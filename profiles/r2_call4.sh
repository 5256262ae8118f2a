mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2_gpu_tests.log 2>&1
tail -15 gpurun_out/r2_gpu_tests.log
BSIG_FUSED_WGRAD=0 timeout 300 python profiles/time_breakdown.py > gpurun_out/r2_breakdown_unfused.txt 2>&1
timeout 300 python profiles/time_breakdown.py > gpurun_out/r2_breakdown_fused.txt 2>&1
grep -E "graph.replay|launches|model.run" gpurun_out/r2_breakdown_unfused.txt gpurun_out/r2_breakdown_fused.txt
