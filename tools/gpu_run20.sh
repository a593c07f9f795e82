#!/bin/bash
# round 2, GPU run 20: two / three contexts (streams) on ONE GPU, each refining a slice of the hypotheses: do the
# slices' kernels fill each other's tail waves?
mkdir -p gpurun_out
for d in 0 0,0 0,0,0 0 0,0; do
  timeout 300 python tools/group_bench.py --devices $d --steps 20 2>&1 | tail -1 | tee -a gpurun_out/r02_two_stream_probe.log
done
