#!/bin/bash
# round 2, GPU run 15: fork threshold 128 + track uploads outside the graph — parity subset, track latency, bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_nets_gpu.py tests/test_pipeline_gpu.py tests/test_register_golden_gpu.py tests/test_track_golden_gpu.py tests/test_multigpu_gpu.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02_pytest15.log
cat gpurun_out/r02_pytest15.log
timeout 300 python tools/bench_track.py 2>&1 | tail -1 | tee gpurun_out/r02_track15.log
timeout 600 python bench.py --steps 20 --warmup 3 2>gpurun_out/r02_bench15.err | tee gpurun_out/r02_bench15.json
