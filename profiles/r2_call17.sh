mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2b_shadowhand_launches2.csv python profiles/shadowhand_step.py 4 > gpurun_out/r2b_sh.log 2>&1; tail -2 gpurun_out/r2b_sh.log
python profiles/summarize_launches.py gpurun_out/r2b_shadowhand_launches2.csv 2>&1 | head -16
