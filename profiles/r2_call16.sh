mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_corr_layer.py -m gpu -q > gpurun_out/r2b_corr_tests.log 2>&1
tail -25 gpurun_out/r2b_corr_tests.log
timeout 600 python - > gpurun_out/r2b_sh_extra.txt 2>&1 <<'PY'
import json, torch, bench
dev = torch.device('cuda', 0)
torch.cuda.set_device(dev)
print(json.dumps(bench.extra_shadowhand(dev), indent=1))
PY
grep -E "ms_per_update|fit_traj|Error|error" gpurun_out/r2b_sh_extra.txt | head
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2b_shadowhand_launches2.csv python profiles/shadowhand_step.py 4 > gpurun_out/r2b_sh.log 2>&1; tail -2 gpurun_out/r2b_sh.log
python profiles/summarize_launches.py gpurun_out/r2b_shadowhand_launches2.csv 2>&1 | head -16
