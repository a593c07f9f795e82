#!/bin/bash
# round 2, GPU run 24/30: the reference's drivers (run_demo.py, run_linemod.py, run_ycb_video.py) unmodified on the drop-in tree
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_dropin_gpu.py -m gpu -x -q -s 2>&1 | tail -30 > gpurun_out/r02_pytest24.log
cat gpurun_out/r02_pytest24.log
