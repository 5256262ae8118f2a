mkdir -p gpurun_out
# shipped f1 kernels: launch list of the ShadowHand update + full ncu capture + DRAM traffic
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file gpurun_out/r2c_shadowhand_launches.csv python profiles/shadowhand_step.py 4 > gpurun_out/r2c_sh.log 2>&1; tail -1 gpurun_out/r2c_sh.log
python profiles/summarize_launches.py gpurun_out/r2c_shadowhand_launches.csv > gpurun_out/r2c_shadowhand_launches.summary.txt 2>&1; head -12 gpurun_out/r2c_shadowhand_launches.summary.txt
ncu --set full --import-source on --clock-control none -k regex:"corr_fwd|corr_wgrad" -s 4 -c 3 -f -o gpurun_out/prof_corr_shipped_r2 python profiles/shadowhand_step.py 4 > /dev/null 2>&1
ls -la gpurun_out/prof_corr_shipped_r2.ncu-rep
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2c_bench_1gpu.json 2> gpurun_out/r2c_bench_1gpu.err
tail -c 300 gpurun_out/r2c_bench_1gpu.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2c_bench_1gpu.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches')}, d['e2e']['value'])
for k, v in d['roofline']['kernels'].items():
    print('%-40s %8.1f GB/s frac %.4f %.4f ms' % (k, v['achieved'], v['frac'], v['ms']))
print(json.dumps(d['extra']['shadowhand_corrdiff_mdnn_1k'])[:600])
PY
