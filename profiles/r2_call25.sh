mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2b_gpu_tests.log 2>&1
tail -8 gpurun_out/r2b_gpu_tests.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2b_bench_1gpu.json 2> gpurun_out/r2b_bench_1gpu.err
tail -c 400 gpurun_out/r2b_bench_1gpu.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2b_bench_1gpu.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches')}, d['e2e']['value'], d['cpu_baseline']['value'])
for k, v in d['extra'].items():
    print(k, json.dumps(v)[:300])
PY
