mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29546 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r2c_bench_4gpu.json 2> gpurun_out/r2c_bench_4gpu.err
tail -c 400 gpurun_out/r2c_bench_4gpu.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2c_bench_4gpu.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'n_gpus')}, d['e2e']['value'], d.get('dp_check'))
print(json.dumps(d.get('extra', {}), indent=0)[:2500])
PY
