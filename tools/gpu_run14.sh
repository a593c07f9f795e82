#!/bin/bash
# round 2, GPU run 14: decoder heads on two streams (fork/join inside the graph) — parity + cost per shard size
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_nets_gpu.py tests/test_pipeline_gpu.py tests/test_register_golden_gpu.py tests/test_track_golden_gpu.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02_pytest14.log
cat gpurun_out/r02_pytest14.log
for f in 0 1000; do
  echo "== FPOSE_FORK_MAX_N=$f" | tee -a gpurun_out/r02_fork_probe.log
  FPOSE_FORK_MAX_N=$f timeout 300 python tools/shard_probe.py 2>&1 | grep shard | tee -a gpurun_out/r02_fork_probe.log
  FPOSE_FORK_MAX_N=$f timeout 300 python tools/bench_track.py 2>&1 | tail -2 | tee -a gpurun_out/r02_fork_probe.log
done
