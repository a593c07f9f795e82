#!/bin/bash
# round 2, GPU run 18: where do the operands of a tracked frame come from? (DRAM bytes / L2 hit rate per kernel, warm caches)
mkdir -p gpurun_out
TRACK_PROFILE=1 timeout 600 ncu --profile-from-start off --cache-control none --clock-control none \
  --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum \
  --csv --log-file gpurun_out/r02_track_cache.csv python tools/bench_track.py > gpurun_out/r02_track_cache.log 2>&1
tail -3 gpurun_out/r02_track_cache.log; wc -l gpurun_out/r02_track_cache.csv
