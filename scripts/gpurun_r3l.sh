#!/bin/bash
mkdir -p gpurun_out
P="python -m pytest -q -p no:cacheprovider"
DRB_MARCH_CELLMAJOR=2 timeout 900 $P tests/test_extract_gpu.py -q -x > gpurun_out/tests_cm5.log 2>&1; echo "extract tests (cm level 5) rc=$?"
for CM in 1 2; do
DRB_MARCH_CELLMAJOR=$CM timeout 600 python bench.py --steps 20 --warmup 3 --streams 1 --no-cpu-baseline > gpurun_out/bench_full_cm$CM.json 2> /dev/null; echo "full cm=$CM rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_full_cm$CM.json').read().strip().splitlines()[-1])
print('cm=$CM', round(d['value'],2), 'ms/step', round(d['ms_per_step'],2), 'marcher ms/launch', round(d['roofline']['kernel_ms_per_launch'],3))
PY
done
