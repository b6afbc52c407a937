#!/bin/bash
# GPU suite + the bench line on the current library (one short GPU call); outputs in gpurun_out/
mkdir -p gpurun_out
T0=$(date +%s)
timeout 300 python -m pytest tests -m gpu -q -n 3 -p no:cacheprovider > gpurun_out/quick_gputests.log 2>&1
echo "pytest rc=$? t=$(( $(date +%s) - T0 ))s"; tail -n 15 gpurun_out/quick_gputests.log | cut -c1-300
timeout 200 python bench.py ${BENCH_ARGS} > gpurun_out/quick_bench_n1.json 2> gpurun_out/quick_bench_n1.err
echo "bench rc=$? t=$(( $(date +%s) - T0 ))s"; tail -n 3 gpurun_out/quick_bench_n1.err | cut -c1-400
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/quick_bench_n1.json").read().strip().splitlines()[-1])
    e = d["e2e"]
    print("value %.0f lat p50 %.1f warm %.1f | e2e %.0f (%.1f us) split_search %.0f sync_search %.0f snapshot %.0f pageable %.0f registered %.0f serial %.0f" % (
        d["value"], d["latency"]["p50"] * 1e3, d["latency"].get("warm_p50", 0) * 1e3, e["value"], e["ms_per_step"] * 1e3,
        e.get("split_search", {}).get("value", 0), e.get("sync_search", {}).get("value", 0), e["snapshot"]["value"], e["pageable_images"]["value"], e["registered_images"]["value"], e["serial_value"]))
except Exception as ex:
    print("no bench line:", ex)
PY
