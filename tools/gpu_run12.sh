#!/bin/bash
mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x > gpurun_out/r02_pytest12_gemm.log 2>&1; echo "gemm rc=$?" >> gpurun_out/r02_pytest12_gemm.log
tail -5 gpurun_out/r02_pytest12_gemm.log
if grep -q "gemm rc=0" gpurun_out/r02_pytest12_gemm.log; then
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/r02_pytest12.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest12.log
timeout 300 python tools/bench_track.py 1000 > gpurun_out/r02_track12.json 2> gpurun_out/r02_track12.err
FPOSE_SPLITK=0 timeout 300 python tools/bench_track.py 1000 > gpurun_out/r02_track12_nosplit.json 2> gpurun_out/r02_track12_nosplit.err
FP_PROFILE_HYP=1 FP_PROFILE_ITERS=2 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_n1_splitk.csv python tools/profile_step.py > gpurun_out/r02_launches_n1_splitk.log 2>&1
grep -E "passed|failed|FAILED|ERROR|worst|track_one over" gpurun_out/r02_pytest12.log | head; cat gpurun_out/r02_track12.json gpurun_out/r02_track12_nosplit.json
fi
