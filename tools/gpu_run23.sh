#!/bin/bash
# round 2, GPU run 23: token reduction as cluster-of-4 (<= 74 sequences) or single CTA (above), same summation order
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_nets_gpu.py tests/test_pipeline_gpu.py tests/test_register_golden_gpu.py tests/test_track_golden_gpu.py tests/test_multigpu_gpu.py -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r02_pytest23.log
cat gpurun_out/r02_pytest23.log
timeout 300 python tools/shard_probe.py 2>&1 | grep shard | tee gpurun_out/r02_shard_probe23.log
timeout 300 python tools/bench_track.py 2>&1 | tail -1 | tee gpurun_out/r02_track23.log
