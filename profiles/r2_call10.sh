mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_mdn.py -m gpu -q > gpurun_out/r2_gemm_tests.log 2>&1
tail -5 gpurun_out/r2_gemm_tests.log
timeout 900 python - > gpurun_out/r2_extras.txt 2>&1 <<'PY'
import json, torch, bench
dev = torch.device('cuda', 0)
torch.cuda.set_device(dev)
print(json.dumps(bench.extra_shadowhand(dev), indent=1))
print(json.dumps(bench.extra_scaled_mode(dev), indent=1))
PY
grep -E "ms_per_update|speedup|fit_traj" gpurun_out/r2_extras.txt
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file gpurun_out/r2_scaled_launches.csv python profiles/scaled_step.py 16384 6 > /dev/null 2>&1; python profiles/summarize_launches.py gpurun_out/r2_scaled_launches.csv 2>&1 | head -14
timeout 600 python -m pytest tests/test_gpu_summarizers.py -m gpu -q > gpurun_out/r2_summ_tests.log 2>&1; tail -6 gpurun_out/r2_summ_tests.log
