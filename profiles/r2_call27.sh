mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_data_parallel.py -m gpu -q > gpurun_out/r2c_dp_tests.log 2>&1; tail -4 gpurun_out/r2c_dp_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c_smoke.log 2>&1; tail -3 gpurun_out/r2c_smoke.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2c_bench_2gpu.json 2> gpurun_out/r2c_bench_2gpu.err
tail -c 500 gpurun_out/r2c_bench_2gpu.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2c_bench_2gpu.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'n_gpus')}, d['e2e']['value'])
print(json.dumps(d.get('extra', {}), indent=0)[:3000])
PY
