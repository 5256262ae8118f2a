mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2d_gpu_tests.log 2>&1
tail -5 gpurun_out/r2d_gpu_tests.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2d_bench_1gpu.json 2> gpurun_out/r2d_bench_1gpu.err
tail -c 300 gpurun_out/r2d_bench_1gpu.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2d_bench_1gpu.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches')}, d['e2e']['value'])
PY
BSIG_FUSED_L0_ADAM=0 timeout 900 python bench.py --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('unfused adam:', d['value'], d['ms_per_step'])"
