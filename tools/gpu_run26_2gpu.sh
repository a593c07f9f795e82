#!/bin/bash
# round 2, GPU run 26 (2 GPUs): multi-GPU tests + 2-GPU bench line of the round-end code (two-stream heads active at 126 per rank)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu_gpu.py -m gpu -q -s > gpurun_out/r02_pytest_2gpu_b.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_2gpu_b.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 --no-standin --no-track > gpurun_out/r02_bench_n2_b.json 2> gpurun_out/r02_bench_n2_b.err
tail -4 gpurun_out/r02_pytest_2gpu_b.log; head -c 600 gpurun_out/r02_bench_n2_b.json; echo; tail -2 gpurun_out/r02_bench_n2_b.err
