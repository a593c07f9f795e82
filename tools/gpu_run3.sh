#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r02_pytest3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest3.log
timeout 600 python tools/score_probe.py > gpurun_out/r02_score_probe.log 2>&1
PROBE_TILES=0,32 timeout 300 python tools/crop_probe.py > gpurun_out/r02c_crop_probe.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench3.json 2> gpurun_out/r02_bench3.err
PROBE_TILES=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:crop_tile -s 2 -c 1 -o gpurun_out/r02c_prof_crop_tile -f python tools/crop_probe.py > gpurun_out/r02c_ncu_crop.log 2>&1
grep -E "passed|failed|FAILED|ERROR|worst|free-running|scorer features|scores:|top-2|driver:|track_one over" gpurun_out/r02_pytest3.log | head -40; cat gpurun_out/r02_score_probe.log; grep "N=" gpurun_out/r02c_crop_probe.log; python -c "
import json;d=json.load(open('gpurun_out/r02_bench3.json'));print({k:d[k] for k in ('value','ms_per_step','best_index','top2_margin','score_spread')});print(d.get('track_one'));print(d.get('gpu_standin'));print(d.get('parity'));print(d['e2e']);print(d['roofline_raster'])"
