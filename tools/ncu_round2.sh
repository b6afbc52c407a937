#!/bin/bash
# One GPU call: launch list + one full capture of the frame workload (run under gpurun; outputs in gpurun_out/).
set -x
FT_PROF_WARMUP=3 FT_PROF_STEPS=8 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ --csv \
  --log-file gpurun_out/r2_launches.csv python tools/profile_frame.py > gpurun_out/r2_prof1.log 2>&1
FT_PROF_WARMUP=2 FT_PROF_STEPS=2 ncu --set full --metrics lts__t_bytes.sum,lts__t_sectors.sum,lts__t_sector_hit_rate.pct \
  --clock-control none --import-source on -k regex:^k_ -s 86 -c 60 -f -o gpurun_out/r2_full python tools/profile_frame.py > gpurun_out/r2_prof2.log 2>&1
ls -la gpurun_out/r2_full.ncu-rep gpurun_out/r2_launches.csv
tail -3 gpurun_out/r2_prof1.log gpurun_out/r2_prof2.log
