#!/bin/bash
FT_PROF_WARMUP=2 FT_PROF_STEPS=1 ncu --section SourceCounters --section WarpStateStats --warp-sampling-interval 0 --import-source on --clock-control none \
  -k regex:'^k_octree' -s 18 -c 1 -f -o /tmp/r2_oct python tools/profile_frame.py > gpurun_out/r2_oct.log 2>&1
ncu -i /tmp/r2_oct.ncu-rep --page source --csv --print-source sass > gpurun_out/r2_oct_source.csv 2>/dev/null
ncu -i /tmp/r2_oct.ncu-rep --page source --csv --print-source cuda > gpurun_out/r2_oct_cuda.csv 2>/dev/null
gzip -f gpurun_out/r2_oct_source.csv; gzip -f gpurun_out/r2_oct_cuda.csv
ls -la gpurun_out/
