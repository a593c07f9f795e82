#!/bin/bash
# round 2, GPU run 33: token reduction over clusters of eight — parity + latency at one pose / 32 hypotheses
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_nets_gpu.py tests/test_pipeline_gpu.py tests/test_register_golden_gpu.py tests/test_track_golden_gpu.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r02_pytest33.log
cat gpurun_out/r02_pytest33.log
timeout 300 python tools/bench_track.py 2>&1 | tail -1 | cut -c 95-260 | tee gpurun_out/r02_track33.log
timeout 300 python tools/shard_probe.py 2>&1 | grep shard | tee gpurun_out/r02_shard_probe33.log
