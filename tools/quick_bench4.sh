#!/bin/bash
mkdir -p gpurun_out
timeout 30 python bench.py --steps 20 --warmup 5 --no-configs --no-cpu-baseline > gpurun_out/qb4.json 2> gpurun_out/qb4.err
echo "rc=$?"; tail -n 2 gpurun_out/qb4.err | cut -c1-300
python -c "
import json
d=json.loads(open('gpurun_out/qb4.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['config'], d['config_detail'])"
