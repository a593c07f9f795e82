#!/bin/bash
# round 2, GPU run 16: phase trace of the tile GEMM launches at 1 and 32 hypotheses
mkdir -p gpurun_out
export FPOSE_LIB_PATH=$PWD/foundationpose_b200/lib/variants/libfpose_trace.so
timeout 300 python tools/gemm_trace.py 1 2 > gpurun_out/r02_gemm_trace_n1.log 2>&1
timeout 300 python tools/gemm_trace.py 32 1 > gpurun_out/r02_gemm_trace_n32.log 2>&1
head -30 gpurun_out/r02_gemm_trace_n1.log
