#!/bin/bash
mkdir -p gpurun_out
PROBE_TILES=0,40 timeout 300 python tools/crop_probe.py > gpurun_out/r02e_crop_probe.log 2>&1
PROBE_TILES=0,40,32 FPOSE_LIB_PATH=$PWD/foundationpose_b200/lib/variants/libfpose_ctas4.so timeout 300 python tools/crop_probe.py > gpurun_out/r02e_crop_probe_ctas4.log 2>&1
timeout 300 python tools/shard_probe.py > gpurun_out/r02_shard_probe.log 2>&1
FP_PROFILE_HYP=32 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_n32.csv python tools/profile_step.py > gpurun_out/r02_launches_n32.log 2>&1
FP_PROFILE_HYP=1 FP_PROFILE_ITERS=2 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_n1.csv python tools/profile_step.py > gpurun_out/r02_launches_n1.log 2>&1
timeout 600 python -m pytest tests/test_raster_gpu.py tests/test_crops_gpu.py -m gpu -q > gpurun_out/r02_pytest7.log 2>&1
grep "N=" gpurun_out/r02e_crop_probe.log gpurun_out/r02e_crop_probe_ctas4.log | grep "mode=0\|---"; cat gpurun_out/r02_shard_probe.log; tail -3 gpurun_out/r02_pytest7.log
