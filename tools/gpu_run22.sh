#!/bin/bash
# round 2, GPU run 22: end-of-round evidence — full GPU suite, smoke, bench line, launch list, ncu --set full of the
# two dominant tile-GEMM shapes after the epilogue change
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r02_pytest22.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest22.log
grep -E "passed|failed|FAILED|ERROR|skipped|worst|free-running|scorer features|scores:|track_one over" gpurun_out/r02_pytest22.log | head -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_smoke22.log 2>&1; tail -2 gpurun_out/r02_smoke22.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench22.json 2> gpurun_out/r02_bench22.err
python -c "
import json;d=json.load(open('gpurun_out/r02_bench22.json'));print({k:d[k] for k in ('value','ms_per_step','best_index','gpu_launches')});print(d['e2e']['value'], d['track_one']['ms_p50'], d['clocks'], d['roofline']['frac'], d['roofline_raster']['avg_launch_ms'])"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches22.csv python tools/profile_step.py > gpurun_out/r02_launches22.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_tile_kernel -s 0 -c 1 -o gpurun_out/r02_prof_conv256_patch -f python tools/profile_step.py > gpurun_out/r02_ncu_conv256.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_tile_kernel -s 6 -c 1 -o gpurun_out/r02_prof_conv512_res -f python tools/profile_step.py > gpurun_out/r02_ncu_conv512.log 2>&1
ls -la gpurun_out/*.ncu-rep
