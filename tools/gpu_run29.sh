#!/bin/bash
# round 2, GPU run 29: why does the e2e - value gap move between 0.1 and 3 ms from box to box?
mkdir -p gpurun_out
timeout 600 python tools/e2e_probe.py > gpurun_out/r02_e2e_probe29.log 2>&1
cat gpurun_out/r02_e2e_probe29.log
