#!/bin/bash
# Source-level (SASS + lineinfo) hot spots of the latency-chain kernels: one launch each, exported to csv on the box.
FT_PROF_WARMUP=2 FT_PROF_STEPS=1 ncu --set full --import-source on --clock-control none \
  -k regex:'^k_(octree|resolve|stereo_match|gather|orient_desc|fast_cells)' -s 44 -c 22 -f -o /tmp/r2_src python tools/profile_frame.py > gpurun_out/r2_src.log 2>&1
ncu -i /tmp/r2_src.ncu-rep --page raw --csv > gpurun_out/r2_src_raw.csv 2>/dev/null
ncu -i /tmp/r2_src.ncu-rep --page source --csv --print-source sass > gpurun_out/r2_src_source.csv 2>/dev/null
ls -la gpurun_out/ /tmp/r2_src.ncu-rep
gzip -f gpurun_out/r2_src_source.csv
ls -la gpurun_out/
