#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r02_pytest10.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest10.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke10.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench10.json 2> gpurun_out/r02_bench10.err
timeout 300 python tools/shard_probe.py > gpurun_out/r02_shard_probe10.log 2>&1
grep -E "passed|failed|FAILED|ERROR|worst|free-running|scorer features|scores:|score spread|track_one over" gpurun_out/r02_pytest10.log | head -30; tail -2 gpurun_out/r02_smoke10.log; python -c "
import json;d=json.load(open('gpurun_out/r02_bench10.json'));print({k:d[k] for k in ('value','ms_per_step','best_index','top2_margin','score_spread','gpu_launches')});print(d['track_one']['ms_p50'],d['track_one']['ms_p99']);print(d['gpu_standin']['value'], d['gpu_standin']['best_index']);print(d['e2e']['value'], d['e2e']['ms_per_step']);print(d['clocks'])"; cat gpurun_out/r02_shard_probe10.log
