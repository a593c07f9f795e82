#!/bin/bash
# 2-GPU pass: NCCL sharded register, fp_group (one process), replica pool; 2-GPU bench lines
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_2gpu_devices.txt 2>&1
timeout 900 python -m pytest tests/test_multigpu_gpu.py -m gpu -q -s > gpurun_out/r02_pytest_2gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err
timeout 600 python tools/replica_bench.py --gpus 2 --frames 32 > gpurun_out/r02_replicas_n2.json 2> gpurun_out/r02_replicas_n2.err
timeout 600 python tools/replica_bench.py --gpus 1 --frames 16 > gpurun_out/r02_replicas_n1.json 2> gpurun_out/r02_replicas_n1.err
tail -5 gpurun_out/r02_pytest_2gpu.log; head -c 700 gpurun_out/r02_bench_n2.json; echo; cat gpurun_out/r02_replicas_n2.json gpurun_out/r02_replicas_n1.json; tail -3 gpurun_out/r02_replicas_n2.err
