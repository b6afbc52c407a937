#!/bin/bash
# per-source-line instruction counts of one launch of the heavy kernels; the (small) report is read locally with tools/ncu_lines.py
FT_PROF_WARMUP=2 FT_PROF_STEPS=1 ncu --section SourceCounters --import-source on --clock-control none \
  -k regex:'^k_(fast_cells|gather|blur|stereo_match|orient_desc)' -s 30 -c 14 -f -o gpurun_out/r2_lines python tools/profile_frame.py > gpurun_out/r2_lines.log 2>&1
ls -la gpurun_out/
