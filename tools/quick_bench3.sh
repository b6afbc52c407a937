#!/bin/bash
mkdir -p gpurun_out
show() { python - $1 <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
e = d["e2e"]
print("value %.0f (%d regions) lat p50 %.1f | e2e %.0f (%.1f us, depth %d) split %.0f sync %.0f snapshot %.0f registered %.0f pageable %.0f" % (
    d["value"], d["timed_regions"], d["latency"]["p50"] * 1e3, e["value"], e["ms_per_step"] * 1e3, e["frames_in_flight"], e["split_search"]["value"], e["sync_search"]["value"],
    e["snapshot"]["value"], e["registered_images"]["value"], e["pageable_images"]["value"]))
PY
}
timeout 120 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/final2_bench_n1_steps20.json 2> gpurun_out/final2_bench_n1_steps20.err
echo "driver command rc=$?"; tail -n 2 gpurun_out/final2_bench_n1_steps20.err | cut -c1-300; show gpurun_out/final2_bench_n1_steps20.json
timeout 60 python bench.py --steps 20 --warmup 5 --no-configs --no-cpu-baseline --pipeline-depth 5 --e2e-depth 5 > gpurun_out/qb_p5_steps20.json 2> gpurun_out/qb_p5.err
echo "depth 5 rc=$?"; show gpurun_out/qb_p5_steps20.json
