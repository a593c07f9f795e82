#!/bin/bash
# round 2, GPU run 28: final state — full GPU suite, smoke, bench line
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -s > gpurun_out/r02_pytest28.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest28.log
grep -E "passed|failed|FAILED|ERROR|skipped|worst|free-running|scorer features|scores:|track_one over|driver:" gpurun_out/r02_pytest28.log | head -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_smoke28.log 2>&1; tail -2 gpurun_out/r02_smoke28.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench28.json 2> gpurun_out/r02_bench28.err
python -c "
import json;d=json.load(open('gpurun_out/r02_bench28.json'));print({k:d[k] for k in ('value','ms_per_step','best_index','gpu_launches')});print(d['e2e']['value'], d['track_one']['ms_p50'], d['clocks'], d['roofline']['frac'], d['roofline_raster']['avg_launch_ms'])"
