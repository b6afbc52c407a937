#!/bin/bash
# A/B run of prebuilt library variants (variants/*.so, git-ignored) on one GPU box: swaps the library in place and prints the
# bench's key numbers for each. Usage: tools/ab_variants.sh v0 v1 ...
set -e
LIB=fasttrack_b200/_build/libfasttrack_b200.so
cp $LIB /tmp/lib_orig.so
for v in "$@"; do
  cp variants/$v.so $LIB
  python bench.py --steps 400 --warmup 10 --no-cpu-baseline > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err || { echo "$v failed"; tail -3 gpurun_out/ab_$v.err; continue; }
  python - "$v" <<'PY'
import json, sys
d = json.load(open("gpurun_out/ab_%s.json" % sys.argv[1]))
s = d["stages_ms"]
print("%-4s value %.0f fps (%.1f us)  latency p50 %.1f us  e2e %.1f us  store %.1f us  gather %.1f resolve %.1f stereo %.1f" % (
    sys.argv[1], d["value"], d["ms_per_step"] * 1e3, d["latency"]["p50"] * 1e3, d["e2e"]["ms_per_step"] * 1e3,
    d["e2e"]["map_store"]["ms_per_step"] * 1e3, s["gather"] * 1e3, s["resolve"] * 1e3, s["stereo_match"] * 1e3))
PY
done
cp /tmp/lib_orig.so $LIB
