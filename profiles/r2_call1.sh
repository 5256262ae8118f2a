mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_call1_smi.txt
./profiles/microbench/mb > gpurun_out/r2_microbench.txt 2>&1
python profiles/time_breakdown.py > gpurun_out/r2_breakdown_default.txt 2>&1
BSIG_CHAIN_TEST=1 timeout 600 python -m pytest tests/test_gpu_chain.py -m gpu -q > gpurun_out/r2_chain_plain.log 2>&1
BSIG_CHAIN=1 timeout 300 python profiles/time_breakdown.py > gpurun_out/r2_breakdown_chain.txt 2>&1
BSIG_CHAIN_TEST=1 timeout 900 compute-sanitizer --tool memcheck --log-file gpurun_out/r2_chain_memcheck.san python -m pytest tests/test_gpu_chain.py -m gpu -q -k "diag or full_big" > gpurun_out/r2_chain_memcheck.log 2>&1
BSIG_CHAIN_TEST=1 timeout 900 compute-sanitizer --tool racecheck --log-file gpurun_out/r2_chain_racecheck.san python -m pytest tests/test_gpu_chain.py -m gpu -q -k "diag" > gpurun_out/r2_chain_racecheck.log 2>&1
tail -3 gpurun_out/r2_chain_plain.log gpurun_out/r2_chain_memcheck.log gpurun_out/r2_chain_racecheck.log
cat gpurun_out/r2_microbench.txt
