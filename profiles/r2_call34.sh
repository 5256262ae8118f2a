mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_final_gpu_tests.log 2>&1
tail -4 gpurun_out/r2_final_gpu_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final_smoke.log 2>&1; tail -1 gpurun_out/r2_final_smoke.log | cut -c1-160
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_final_bench_1gpu.json 2> gpurun_out/r2_final_bench_1gpu.err
tail -c 300 gpurun_out/r2_final_bench_1gpu.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_final_bench_reference.json 2> gpurun_out/r2_final_bench_reference.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_final_bench_1gpu.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches')}, 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['kind'])
print('clocks', d.get('clocks'))
for k, v in d['roofline']['kernels'].items():
    print('%-40s %8.1f GB/s frac %.4f %.4f ms' % (k, v['achieved'], v['frac'], v['ms']))
print(json.dumps(d['extra']['shadowhand_corrdiff_mdnn_1k'])[-330:])
r = json.loads(open('gpurun_out/r2_final_bench_reference.json').read().strip().splitlines()[-1])
print('reference arm:', {k: r.get(k) for k in ('impl', 'value', 'unit', 'ms_per_step')}, r['cpu_baseline']['kind'])
PY
