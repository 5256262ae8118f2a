mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none -k regex:"corr_wgrad" -s 2 -c 1 -f -o gpurun_out/prof_corr_wg_r2 python profiles/shadowhand_step.py 4 > gpurun_out/r2b_ncu_corr.log 2>&1
tail -2 gpurun_out/r2b_ncu_corr.log; ls -la gpurun_out/prof_corr_wg_r2.ncu-rep
