#!/bin/bash
# compute-sanitizer over a few frames of the bench workload (core path), both octree paths
export FT_PROF_WARMUP=1 FT_PROF_STEPS=2
for tool in memcheck racecheck; do
  for dense in 1 0; do
    FT_OCT_DENSE=$dense timeout 600 compute-sanitizer --tool $tool --print-limit 20 python tools/profile_frame.py > gpurun_out/r2_sanitizer_${tool}_dense${dense}.log 2>&1
    echo "$tool dense=$dense: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r2_sanitizer_${tool}_dense${dense}.log | tail -1)"
  done
done
