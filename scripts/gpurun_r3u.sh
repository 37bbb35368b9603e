#!/bin/bash
mkdir -p gpurun_out
nproc
for B in 1 0 1; do
DRB_BLOCKING_SYNC=$B timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_default_x.json 2> /dev/null
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_default_x.json').read().strip().splitlines()[-1])
print('blocking=$B: value', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), d['config'].get('host_wait'), 'one pair at a time', d['config'].get('streams','')[-22:])
PY
done
DRB_BLOCKING_SYNC=1 timeout 600 python bench.py --stage batch --pairs-per-gpu 32 --steps 3 --warmup 3 > gpurun_out/bench_batch_x.json 2> /dev/null
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_batch_x.json').read().strip().splitlines()[-1])
print('batch blocking: value', round(d['value'],2), 'e2e', round(d['e2e']['value'],2))
PY
