#!/bin/bash
# round 2, GPU run 27 (8 GPUs): 8- and 4-GPU bench lines of the round-end code
mkdir -p gpurun_out
for n in 8 4; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 3 --no-standin --no-track > gpurun_out/r02_bench_n${n}_b.json 2> gpurun_out/r02_bench_n${n}_b.err
  head -c 420 gpurun_out/r02_bench_n${n}_b.json; echo
done
