#!/bin/bash
mkdir -p gpurun_out
nvidia-smi topo -m 2>/dev/null | head -12
timeout 120 python scripts/h2d_numa.py
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_default_1.json 2> /dev/null; echo "rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_default_1.json').read().strip().splitlines()[-1])
print('default', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), d['e2e']['ms_per_step'])
PY
