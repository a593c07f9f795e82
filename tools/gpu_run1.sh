#!/bin/bash
# first GPU pass of round 2: tests, crop probe (+ register-allocation variants), net probe, short bench, ncu of the crop kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/r02_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/r02_pytest1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest1.log
timeout 300 python tools/crop_probe.py > gpurun_out/r02_crop_probe.log 2>&1
for c in 2 4; do FPOSE_LIB_PATH=$PWD/foundationpose_b200/lib/variants/libfpose_ctas$c.so timeout 300 python tools/crop_probe.py > gpurun_out/r02_crop_probe_ctas$c.log 2>&1; done
timeout 300 python tools/net_probe.py > gpurun_out/r02_net_probe.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench1.json 2> gpurun_out/r02_bench1.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:crop_tile -s 2 -c 1 -o gpurun_out/r02_prof_crop_tile -f python tools/crop_probe.py > gpurun_out/r02_ncu_crop.log 2>&1
tail -5 gpurun_out/r02_pytest1.log; cat gpurun_out/r02_crop_probe.log; tail -3 gpurun_out/r02_net_probe.log; cat gpurun_out/r02_bench1.json | head -c 1500
