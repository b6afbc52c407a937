#!/bin/bash
# One GPU call: launch list + one full capture of ONE frame of the workload (run under gpurun; outputs in gpurun_out/, which
# must stay below 64 MiB: the raw metric table is exported to csv on the box and the report is dropped if it is large).
FT_PROF_WARMUP=3 FT_PROF_STEPS=8 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ --csv \
  --log-file gpurun_out/r2_launches.csv python tools/profile_frame.py > gpurun_out/r2_prof1.log 2>&1
FT_PROF_WARMUP=2 FT_PROF_STEPS=2 ncu --set full --metrics lts__t_bytes.sum,lts__t_sectors.sum,lts__t_sector_hit_rate.pct \
  --clock-control none -k regex:^k_ -s 86 -c 30 -f -o gpurun_out/r2_full python tools/profile_frame.py > gpurun_out/r2_prof2.log 2>&1
ncu -i gpurun_out/r2_full.ncu-rep --page raw --csv > gpurun_out/r2_full_raw.csv 2> gpurun_out/r2_export.log
sz=$(stat -c %s gpurun_out/r2_full.ncu-rep)
if [ "$sz" -gt 40000000 ]; then rm gpurun_out/r2_full.ncu-rep; fi
ls -la gpurun_out/
tail -n 3 gpurun_out/r2_prof1.log gpurun_out/r2_prof2.log
