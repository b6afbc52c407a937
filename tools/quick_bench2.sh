#!/bin/bash
mkdir -p gpurun_out
for de in 4 5; do
timeout 100 python bench.py --no-configs --no-cpu-baseline --e2e-depth $de --pipeline-depth $(( de > 4 ? de : 4 )) > gpurun_out/qb_de$de.json 2> gpurun_out/qb_de$de.err
echo "de=$de bench rc=$?"; tail -n 2 gpurun_out/qb_de$de.err | cut -c1-300
python - $de <<'PY'
import json, sys
d = json.loads(open("gpurun_out/qb_de%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
e = d["e2e"]
print("value %.0f lat p50 %.1f | e2e %.0f (%.1f us) split %.0f sync %.0f snapshot %.0f registered %.0f pageable %.0f" % (
    d["value"], d["latency"]["p50"] * 1e3, e["value"], e["ms_per_step"] * 1e3, e["split_search"]["value"], e["sync_search"]["value"],
    e["snapshot"]["value"], e["registered_images"]["value"], e["pageable_images"]["value"]))
print("phases e2e :", {k: round(v, 1) for k, v in e["host_phases_us_per_frame"].items()})
PY
done
