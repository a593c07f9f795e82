#!/bin/bash
mkdir -p gpurun_out
PROBE_OUTLIERS=1 timeout 600 python tools/score_probe.py > gpurun_out/r02_score_outliers.log 2>&1
timeout 300 python tools/e2e_probe.py > gpurun_out/r02_e2e_probe_pin.log 2>&1
FPOSE_NO_PIN=1 timeout 300 python tools/e2e_probe.py > gpurun_out/r02_e2e_probe_nopin.log 2>&1
for v in u2c3 u2c2; do PROBE_TILES=0 FPOSE_LIB_PATH=$PWD/foundationpose_b200/lib/variants/libfpose_$v.so timeout 300 python tools/crop_probe.py > gpurun_out/r02d_crop_probe_$v.log 2>&1; done
timeout 900 python -m pytest tests/test_multigpu_gpu.py tests/test_raster_gpu.py -m gpu -q -s > gpurun_out/r02_pytest4.log 2>&1
cat gpurun_out/r02_score_outliers.log gpurun_out/r02_e2e_probe_pin.log gpurun_out/r02_e2e_probe_nopin.log; grep "N=252\|N= 32 mode=0\|N=  1 mode=0" gpurun_out/r02d_crop_probe_*.log; tail -5 gpurun_out/r02_pytest4.log
