#!/bin/bash
# round 2, GPU run 17: exposed-epilogue fix (bias tile in smem, pos.-emb. prefetch), narrow tiles for tiny grids,
# weights-before-dependency-wait experiment
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_nets_gpu.py tests/test_pipeline_gpu.py tests/test_register_golden_gpu.py tests/test_track_golden_gpu.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02_pytest17.log
cat gpurun_out/r02_pytest17.log
T=$PWD/foundationpose_b200/lib/variants/libfpose_trace.so
FPOSE_LIB_PATH=$T timeout 300 python tools/gemm_trace.py 1 2 > gpurun_out/r02_gemm_trace17_n1.log 2>&1
FPOSE_LIB_PATH=$T FPOSE_GEMM_BPRE=1 timeout 300 python tools/gemm_trace.py 1 2 > gpurun_out/r02_gemm_trace17_n1_bpre.log 2>&1
FPOSE_LIB_PATH=$T FPOSE_NARROW=0 timeout 300 python tools/gemm_trace.py 1 2 > gpurun_out/r02_gemm_trace17_n1_wide.log 2>&1
FPOSE_LIB_PATH=$T timeout 300 python tools/gemm_trace.py 32 1 > gpurun_out/r02_gemm_trace17_n32.log 2>&1
FPOSE_LIB_PATH=$T FPOSE_GEMM_BPRE=1 timeout 300 python tools/gemm_trace.py 32 1 > gpurun_out/r02_gemm_trace17_n32_bpre.log 2>&1
for b in 0 1; do
  echo "== FPOSE_GEMM_BPRE=$b" | tee -a gpurun_out/r02_bpre_probe.log
  FPOSE_GEMM_BPRE=$b timeout 300 python tools/shard_probe.py 2>&1 | grep shard | tee -a gpurun_out/r02_bpre_probe.log
  FPOSE_GEMM_BPRE=$b timeout 300 python tools/bench_track.py 2>&1 | tail -1 | tee -a gpurun_out/r02_bpre_probe.log
done
echo "== FPOSE_NARROW=0" | tee -a gpurun_out/r02_bpre_probe.log
FPOSE_NARROW=0 timeout 300 python tools/bench_track.py 2>&1 | tail -1 | tee -a gpurun_out/r02_bpre_probe.log
