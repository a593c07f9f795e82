#!/bin/bash
# round 2, GPU run 19: what bounds the K loop of a one-tile CTA?  (OOB zero-fill vs gather rate vs MMA)
mkdir -p gpurun_out
export FPOSE_LIB_PATH=$PWD/foundationpose_b200/lib/variants/libfpose_trace.so
for n in 8 4; do timeout 300 python tools/gemm_trace.py $n 1 > gpurun_out/r02_gemm_trace19_n$n.log 2>&1; done
FPOSE_CTA_GROUP=1 timeout 300 python tools/gemm_trace.py 1 1 > gpurun_out/r02_gemm_trace19_n1_cg1.log 2>&1
FPOSE_CTA_GROUP=1 timeout 300 python tools/gemm_trace.py 8 1 > gpurun_out/r02_gemm_trace19_n8_cg1.log 2>&1
head -14 gpurun_out/r02_gemm_trace19_n8.log
