mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py -m gpu -q -x > gpurun_out/r2b_gemm_tests.log 2>&1
tail -5 gpurun_out/r2b_gemm_tests.log
timeout 600 python profiles/rooflines_only.py > gpurun_out/r2b_rooflines.txt 2>&1; cat gpurun_out/r2b_rooflines.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2b_shadowhand_launches.csv python profiles/shadowhand_step.py 4 > gpurun_out/r2b_sh.log 2>&1; tail -2 gpurun_out/r2b_sh.log
python profiles/summarize_launches.py gpurun_out/r2b_shadowhand_launches.csv 2>&1 | head -24
