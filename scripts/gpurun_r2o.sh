#!/bin/bash
mkdir -p gpurun_out
P="python -m pytest -q -p no:cacheprovider"
timeout 900 $P tests/test_backward_gpu.py -x -q > gpurun_out/tests_bwd.log 2>&1; echo "backward tests rc=$?"; tail -n 4 gpurun_out/tests_bwd.log
timeout 900 python bench.py --stage train --batch 32 --steps 2 --warmup 3 --no-cpu-baseline --streams 4 > gpurun_out/bench_train_b32_s4.json 2> gpurun_out/bench_train_b32_s4.err; echo "train rc=$?"
timeout 900 python bench.py --stage train --batch 32 --steps 2 --warmup 3 --no-cpu-baseline --streams 1 > gpurun_out/bench_train_b32_s1.json 2> gpurun_out/bench_train_b32_s1.err; echo "train rc=$?"
for f in train_b32_s4 train_b32_s1; do python - <<PY
import json
d=json.loads(open('gpurun_out/bench_${f}.json').read().strip().splitlines()[-1])
print('${f}', round(d['value'],2), d['unit'], 'e2e', round(d['e2e']['value'],2), 'ms/step', round(d['ms_per_step'],2), 'roofline frac', round(d['roofline']['frac'],4))
PY
done
