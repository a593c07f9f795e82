#!/bin/bash
# round 2, GPU run 31: fused frame preparation (erode + bilateral + xyz + rgba in one launch): parity + track latency
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_crops_gpu.py tests/test_pipeline_gpu.py tests/test_register_golden_gpu.py tests/test_track_golden_gpu.py tests/test_raster_gpu.py -m gpu -x -q -s 2>&1 | tail -12 > gpurun_out/r02_pytest31.log
cat gpurun_out/r02_pytest31.log
timeout 300 python tools/bench_track.py 2>&1 | tail -1 | tee gpurun_out/r02_track31.log
timeout 300 python tools/fixed_cost_probe.py 2>&1 | tail -8 | tee gpurun_out/r02_fixed_cost31.log
