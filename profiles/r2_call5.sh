mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_gpu_tests.log 2>&1
tail -40 gpurun_out/r2_gpu_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -3 gpurun_out/r2_smoke.log
timeout 300 python profiles/time_breakdown.py > gpurun_out/r2_breakdown.txt 2>&1; cat gpurun_out/r2_breakdown.txt | tail -12
