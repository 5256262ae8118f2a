mkdir -p gpurun_out
# DRAM traffic + duration of every shipped streaming kernel at the roofline sizes (metrics-only pass)
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2_roofline_traffic.csv python profiles/rooflines_only.py > gpurun_out/r2_rooflines_under_ncu.log 2>&1
# event-timed numbers (not under a profiler)
python profiles/rooflines_only.py > gpurun_out/r2_rooflines.txt 2>&1; cat gpurun_out/r2_rooflines.txt
# full captures of the shipped tcgen05 GEMM and signature kernels
ncu --set full --import-source on --clock-control none -k regex:gemm_tc_kernel -s 6 -c 2 -o gpurun_out/prof_gemm_tc_r2 python profiles/rooflines_only.py > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:signature3_small -s 3 -c 1 -o gpurun_out/prof_sig_r2 python profiles/rooflines_only.py > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:"nll_stream|exp_sum|eps_fixup" -s 9 -c 3 -o gpurun_out/prof_nll_r2 python profiles/rooflines_only.py > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
