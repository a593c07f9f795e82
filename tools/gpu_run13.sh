#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r02_pytest13.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest13.log
grep -E "passed|failed|FAILED|ERROR|skipped|worst|free-running|scorer features|scores:|track_one over" gpurun_out/r02_pytest13.log | head -30
