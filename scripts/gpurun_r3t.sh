#!/bin/bash
mkdir -p gpurun_out
nproc
for S in 4 4 2 4; do
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --streams $S > gpurun_out/bench_default_x.json 2> /dev/null
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_default_x.json').read().strip().splitlines()[-1])
print('streams $S: value', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'e2e ms/step', round(d['e2e']['ms_per_step'],2))
PY
done
