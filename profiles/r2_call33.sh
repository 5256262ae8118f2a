mkdir -p gpurun_out
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 1 --warmup 3 > gpurun_out/r2_bench_under_ncu.log 2>&1
python profiles/summarize_launches.py gpurun_out/r2_launches_bench.csv > gpurun_out/r2_launches_bench.summary.txt 2>&1; head -14 gpurun_out/r2_launches_bench.summary.txt
gzip -f gpurun_out/r2_launches_bench.csv; ls -la gpurun_out/r2_launches_bench.csv.gz
