#!/bin/bash
# compute-sanitizer over the final library: a frame of the bench workload (memcheck + racecheck) and the projection-search
# tests (memcheck), i.e. the kernels changed last (k_resolve with distributed shared memory, k_gather with bulk copies);
# then the bench with the driver's arguments. Outputs in gpurun_out/.
mkdir -p gpurun_out
T0=$(date +%s)
export FT_PROF_WARMUP=1 FT_PROF_STEPS=2
for tool in memcheck racecheck; do
  timeout 100 compute-sanitizer --tool $tool --print-limit 20 python tools/profile_frame.py > gpurun_out/r2_sanitizer_${tool}_final_frame.log 2>&1
  echo "$tool frame rc=$? t=$(( $(date +%s) - T0 ))s: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r2_sanitizer_${tool}_final_frame.log | tail -1)"
done
timeout 150 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest -q -p no:cacheprovider -m gpu \
  tests/test_gpu_last_frame.py tests/test_gpu_store.py tests/test_gpu_fisheye.py tests/test_gpu_rgbd.py \
  > gpurun_out/r2_sanitizer_memcheck_final_search_tests.log 2>&1
echo "memcheck tests rc=$? t=$(( $(date +%s) - T0 ))s: $(grep -E 'ERROR SUMMARY|passed|failed' gpurun_out/r2_sanitizer_memcheck_final_search_tests.log | tail -2 | tr '\n' ' ')"
timeout 60 python bench.py --steps 20 --warmup 5 --no-configs > gpurun_out/final_bench_n1_steps20.json 2> gpurun_out/final_bench_n1_steps20.err
echo "bench steps20 rc=$? t=$(( $(date +%s) - T0 ))s"
cut -c1-300 gpurun_out/final_bench_n1_steps20.json
