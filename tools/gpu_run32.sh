#!/bin/bash
# round 2, GPU run 32: fused frame preparation A/B on one box
mkdir -p gpurun_out
for rep in 1 2; do for f in 1 0; do
  echo "== FPOSE_FUSED_PREP=$f" | tee -a gpurun_out/r02_prep_ab.log
  FPOSE_FUSED_PREP=$f timeout 300 python tools/bench_track.py 2>&1 | tail -1 | cut -c 95-260 | tee -a gpurun_out/r02_prep_ab.log
done; done
for f in 1 0; do FPOSE_FUSED_PREP=$f timeout 300 python tools/fixed_cost_probe.py 2>&1 | grep set_frame | tee -a gpurun_out/r02_prep_ab.log; done
