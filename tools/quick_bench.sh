#!/bin/bash
# bench line only (short GPU call): BENCH_ARGS default to the fast variant
mkdir -p gpurun_out
timeout 150 python bench.py ${BENCH_ARGS:---no-configs --no-cpu-baseline} > gpurun_out/qb.json 2> gpurun_out/qb.err
echo "bench rc=$?"; tail -n 3 gpurun_out/qb.err | cut -c1-400
python - <<'PY'
import json
d = json.loads(open("gpurun_out/qb.json").read().strip().splitlines()[-1])
e = d["e2e"]
print("value %.0f lat p50 %.1f | e2e %.0f (%.1f us) split %.0f sync %.0f snapshot %.0f registered %.0f" % (
    d["value"], d["latency"]["p50"] * 1e3, e["value"], e["ms_per_step"] * 1e3, e["split_search"]["value"], e["sync_search"]["value"],
    e["snapshot"]["value"], e["registered_images"]["value"]))
print("phases e2e :", {k: round(v, 1) for k, v in e["host_phases_us_per_frame"].items()})
print("phases sync:", {k: round(v, 1) for k, v in e["sync_search"]["host_phases_us_per_frame"].items()})
PY
