#!/bin/bash
# One GPU call on the final code: GPU parity suite, the bench line, the ncu launch list and (time permitting) one full capture
# of a frame. Every leg has its own timeout; outputs in gpurun_out/.
mkdir -p gpurun_out
T0=$(date +%s)
git rev-parse HEAD > gpurun_out/final_head.txt 2>/dev/null
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/final_smi.txt
timeout 300 python -m pytest tests -m gpu -q -n 3 -p no:cacheprovider > gpurun_out/final_gputests.log 2>&1
echo "pytest rc=$? t=$(( $(date +%s) - T0 ))s" | tee -a gpurun_out/final_gputests.log
tail -n 5 gpurun_out/final_gputests.log
timeout 240 python bench.py > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err
echo "bench rc=$? t=$(( $(date +%s) - T0 ))s"
cut -c1-600 gpurun_out/final_bench_n1.json
FT_PROF_WARMUP=3 FT_PROF_STEPS=8 timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ --csv \
  --log-file gpurun_out/r2_launches.csv python tools/profile_frame.py > gpurun_out/r2_prof1.log 2>&1
echo "launch list rc=$? t=$(( $(date +%s) - T0 ))s"
FT_PROF_WARMUP=2 FT_PROF_STEPS=2 timeout 200 ncu --set full --metrics lts__t_bytes.sum,lts__t_sectors.sum,lts__t_sector_hit_rate.pct \
  --clock-control none -k regex:^k_ -s 86 -c 30 -f -o gpurun_out/r2_full python tools/profile_frame.py > gpurun_out/r2_prof2.log 2>&1
echo "full capture rc=$? t=$(( $(date +%s) - T0 ))s"
timeout 60 ncu -i gpurun_out/r2_full.ncu-rep --page raw --csv > gpurun_out/r2_full_raw.csv 2> gpurun_out/r2_export.log
sz=$(stat -c %s gpurun_out/r2_full.ncu-rep 2>/dev/null || echo 0)
if [ "$sz" -gt 40000000 ]; then rm gpurun_out/r2_full.ncu-rep; fi
ls -la gpurun_out/
