#!/bin/bash
mkdir -p gpurun_out
PROBE_WORST=1 timeout 600 python tools/score_probe.py > gpurun_out/r02_score_worst.log 2>&1
timeout 900 python -m pytest tests/test_nets_gpu.py tests/test_pipeline_gpu.py tests/test_register_golden_gpu.py -m gpu -q -s > gpurun_out/r02_pytest5.log 2>&1
cat gpurun_out/r02_score_worst.log; grep -E "passed|failed|FAILED|scorer features|scores:|top-2" gpurun_out/r02_pytest5.log
