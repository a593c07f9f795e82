#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r02_pytest6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest6.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench6.json 2> gpurun_out/r02_bench6.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench6_ref.json 2> gpurun_out/r02_bench6_ref.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_register_252x5.csv python tools/profile_step.py > gpurun_out/r02_launches.log 2>&1
grep -E "passed|failed|FAILED|ERROR|worst|free-running|scorer features|scores:|score spread|driver:|track_one over" gpurun_out/r02_pytest6.log | head -40; cat gpurun_out/r02_smoke.log | tail -3; python -c "
import json;d=json.load(open('gpurun_out/r02_bench6.json'));print({k:d[k] for k in ('value','ms_per_step','best_index','top2_margin','score_spread','gpu_launches')});print(d.get('track_one'));print(d.get('gpu_standin'));print(d.get('parity'));print(d['e2e']);print(d['clocks']);print(d['roofline']);print(d['roofline_raster']);print(d.get('cpu_baseline'))"; cat gpurun_out/r02_bench6_ref.json | cut -c1-600
