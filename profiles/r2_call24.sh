mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_corr_layer.py -m gpu -q -x > gpurun_out/r2b_corr_tests.log 2>&1
tail -3 gpurun_out/r2b_corr_tests.log
timeout 600 python - > gpurun_out/r2b_sh_extra.txt 2>&1 <<'PY'
import json, os, torch, bench
dev = torch.device('cuda', 0)
torch.cuda.set_device(dev)
print(json.dumps(bench.extra_shadowhand(dev), indent=1))
os.environ['BSIG_FUSED_CORR'] = '0'
print(json.dumps(bench.extra_shadowhand(dev), indent=1))
PY
grep -E "ms_per_update|fit_traj|final_test|Error|error" gpurun_out/r2b_sh_extra.txt | head
