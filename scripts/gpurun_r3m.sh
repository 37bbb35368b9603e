#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --stage register --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_register.json 2> /dev/null; echo "register rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_register.json').read().strip().splitlines()[-1])
print('register', round(d['value'],2), 'ms/step', round(d['ms_per_step'],3), 'igemm ms', round(d['roofline']['kernel_ms_per_step'],3))
PY
