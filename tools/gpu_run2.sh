#!/bin/bash
# second GPU pass of round 2: all GPU tests (no -x), crop probe over tile sizes and register-allocation variants, bench, ncu
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r02_pytest2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest2.log
PROBE_TILES=16,32,80 timeout 300 python tools/crop_probe.py > gpurun_out/r02b_crop_probe.log 2>&1
for c in 2 4; do PROBE_TILES=0 FPOSE_LIB_PATH=$PWD/foundationpose_b200/lib/variants/libfpose_ctas$c.so timeout 300 python tools/crop_probe.py > gpurun_out/r02b_crop_probe_ctas$c.log 2>&1; done
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench2.json 2> gpurun_out/r02_bench2.err
PROBE_TILES=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:crop_tile -s 2 -c 1 -o gpurun_out/r02b_prof_crop_tile -f python tools/crop_probe.py > gpurun_out/r02b_ncu_crop.log 2>&1
grep -E "passed|failed|FAILED|ERROR|worst|free-running|scorer features|scores:|top-2|driver:" gpurun_out/r02_pytest2.log | head -40; cat gpurun_out/r02b_crop_probe.log; grep "N=252 mode=0" gpurun_out/r02b_crop_probe_ctas*.log; head -c 600 gpurun_out/r02_bench2.json
