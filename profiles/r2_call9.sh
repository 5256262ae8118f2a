mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_mdn.py tests/test_gpu_bayessim.py -m gpu -q > gpurun_out/r2_gemm_tests.log 2>&1
tail -12 gpurun_out/r2_gemm_tests.log
timeout 900 python - > gpurun_out/r2_extras.txt 2>&1 <<'PY'
import json, torch, bench
dev = torch.device('cuda', 0)
torch.cuda.set_device(dev)
print(json.dumps(bench.extra_shadowhand(dev), indent=1))
print(json.dumps(bench.extra_scaled_mode(dev), indent=1))
PY
cat gpurun_out/r2_extras.txt | tail -60
