#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_default_$i.json 2> /dev/null; echo "rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_default_$i.json').read().strip().splitlines()[-1])
print('default run $i', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), d['e2e']['ms_per_step'])
PY
done
