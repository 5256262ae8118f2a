mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_corr_layer.py -m gpu -q -x > gpurun_out/r2b_corr_tests.log 2>&1
tail -5 gpurun_out/r2b_corr_tests.log
timeout 300 python profiles/corr_dbg.py 2>&1 | tail -5
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2b_shadowhand_launches2.csv python profiles/shadowhand_step.py 4 > gpurun_out/r2b_sh.log 2>&1; tail -2 gpurun_out/r2b_sh.log
python profiles/summarize_launches.py gpurun_out/r2b_shadowhand_launches2.csv 2>&1 | head -6
