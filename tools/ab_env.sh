#!/bin/bash
# bench.py key numbers under different environment settings / flags. Usage: tools/ab_env.sh "ENV=1 ..|flags" ...
for spec in "$@"; do
  envs="${spec%%|*}"; flags="${spec#*|}"
  env $envs python bench.py --no-configs --no-cpu-baseline $flags > /tmp/ab_out.json 2> /tmp/ab_err.txt
  if [ -s /tmp/ab_out.json ]; then
    python - "$spec" <<'PY'
import json, sys
d = json.loads(open("/tmp/ab_out.json").read().strip().splitlines()[-1])
print("%-60s value %.0f  lat p50 %.1f warm %.1f  e2e %.0f  snapshot %.0f serial %.0f" % (
    sys.argv[1], d["value"], d["latency"]["p50"] * 1e3, d["latency"].get("warm_p50", 0) * 1e3, d["e2e"]["value"],
    d["e2e"]["snapshot"]["value"], d["e2e"]["serial_value"]))
PY
  else
    echo "$spec FAILED: $(tail -2 /tmp/ab_err.txt | cut -c1-300)"
  fi
done
