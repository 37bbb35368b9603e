#!/bin/bash
mkdir -p gpurun_out
P="python -m pytest -q -p no:cacheprovider"
timeout 1200 $P tests/test_backward_gpu.py -q -x > gpurun_out/tests_bwd2.log 2>&1; echo "backward tests rc=$?"; grep -E "^(FAILED|ERROR)" gpurun_out/tests_bwd2.log | head -5
timeout 900 python bench.py --stage train --batch 32 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_train_b32_s4.json 2> /dev/null; echo "train rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_train_b32_s4.json').read().strip().splitlines()[-1])
print('train_b32_s4', round(d['value'],2), d['unit'], 'e2e', round(d['e2e']['value'],2), 'ms/step', round(d['ms_per_step'],2))
PY
