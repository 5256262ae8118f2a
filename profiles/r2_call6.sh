mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_data_parallel.py -m gpu -q > gpurun_out/r2_dp_tests.log 2>&1; tail -5 gpurun_out/r2_dp_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err; tail -c 3000 gpurun_out/r2_bench_2gpu.json; tail -5 gpurun_out/r2_bench_2gpu.err
