#!/bin/bash
# 8-GPU pass: strong scaling of the sharded register (torchrun, NCCL), fp_group on 8 devices, replica pool on 8 devices
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_8gpu_devices.txt 2>&1
for n in 8; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/r02_bench_n$n.json 2> gpurun_out/r02_bench_n$n.err
done
timeout 600 python tools/group_bench.py --gpus 8 > gpurun_out/r02_group_n8.json 2> gpurun_out/r02_group_n8.err
timeout 600 python tools/replica_bench.py --gpus 8 --frames 128 > gpurun_out/r02_replicas_n8.json 2> gpurun_out/r02_replicas_n8.err
for n in 8; do python -c "
import json,sys;d=json.load(open('gpurun_out/r02_bench_n$n.json'));print('$n',{k:d[k] for k in ('value','ms_per_step','best_index','gpu_launches')}, d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['clocks']['sm_mhz'])"; done
cat gpurun_out/r02_group_n8.json gpurun_out/r02_replicas_n8.json; tail -2 gpurun_out/r02_group_n8.err gpurun_out/r02_replicas_n8.err
