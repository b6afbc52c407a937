#!/bin/bash
# A/B run of prebuilt library variants (variants/*.so, git-ignored) on one GPU box: swaps the library in place and prints the
# bench's key numbers for each, interleaved twice. Usage: tools/ab_variants.sh v0 v1 ...
LIB=fasttrack_b200/_build/libfasttrack_b200.so
cp $LIB /tmp/lib_orig.so
for rep in $(seq 1 ${REPS:-2}); do
for v in "$@"; do
  cp variants/$v.so $LIB
  python bench.py --no-configs --no-cpu-baseline > /tmp/ab_out.json 2> /tmp/ab_err.txt
  if [ -s /tmp/ab_out.json ]; then
    python - "$v" <<'PY'
import json, sys
d = json.loads(open("/tmp/ab_out.json").read().strip().splitlines()[-1]); s = d["stages_isolated_ms"]
p = d["stages_ms"]
print("%-8s value %.0f  lat p50 %.1f warm %.1f  e2e %.0f  snapshot %.0f  isolated: gather %.1f resolve %.1f octree_l0 %.1f orient %.1f stereo %.1f  "
      "in the frame: gather %.1f resolve %.1f" % (
    sys.argv[1], d["value"], d["latency"]["p50"] * 1e3, d["latency"].get("warm_p50", 0) * 1e3, d["e2e"]["value"],
    d["e2e"]["snapshot"]["value"], s["gather"] * 1e3, s["resolve"] * 1e3, s["octree_l0"] * 1e3, s["orient_desc"] * 1e3,
    s["stereo_match"] * 1e3, p["gather"] * 1e3, p["resolve"] * 1e3), "fast %.1f+%.1f blur %.1f+%.1f" % (s["fast_cells_l0"] * 1e3, s["fast_cells"] * 1e3, s["blur_l0"] * 1e3, s["blur"] * 1e3))
PY
  else
    echo "$v FAILED: $(tail -2 /tmp/ab_err.txt | cut -c1-300)"
  fi
done
done
cp /tmp/lib_orig.so $LIB
