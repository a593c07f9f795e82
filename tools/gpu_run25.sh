#!/bin/bash
# round 2, GPU run 25: compute-sanitizer over the GPU suite (memcheck: all kernels; racecheck: the shared-memory z-tile /
# binning of the crop producer and the token reductions)
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/r02_memcheck.log \
  python -m pytest tests/test_pipeline_gpu.py tests/test_nets_gpu.py tests/test_crops_gpu.py tests/test_raster_gpu.py tests/test_track_golden_gpu.py tests/test_gemm_gpu.py -m gpu -q > gpurun_out/r02_memcheck_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/r02_memcheck_pytest.log
tail -4 gpurun_out/r02_memcheck_pytest.log; tail -3 gpurun_out/r02_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file gpurun_out/r02_racecheck.log \
  python -m pytest tests/test_crops_gpu.py tests/test_raster_gpu.py -m gpu -q -x > gpurun_out/r02_racecheck_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/r02_racecheck_pytest.log
tail -4 gpurun_out/r02_racecheck_pytest.log; tail -3 gpurun_out/r02_racecheck.log
