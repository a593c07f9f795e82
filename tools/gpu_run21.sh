#!/bin/bash
# round 2, GPU run 21: at the shard sizes of 8 / 4 / 2 GPUs, does splitting the shard over two contexts (streams) on the
# same GPU hide the per-launch bubbles?
mkdir -p gpurun_out
for h in 32 63 126; do
  for d in 0 0,0 0 0,0; do
    timeout 300 python tools/group_bench.py --devices $d --hyp $h --steps 40 2>&1 | tail -1 | cut -c 150- | tee -a gpurun_out/r02_two_stream_shard_probe.log
  done
done
