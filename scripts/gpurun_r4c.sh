#!/bin/bash
# weight gradient: K splits from the wave model - tests, per-shape A/B on one box, training step
mkdir -p gpurun_out
P="python -m pytest -q -p no:cacheprovider"
timeout 600 $P tests/test_backward_gpu.py -m gpu -x > gpurun_out/tests_wg2.log 2>&1; echo "tests rc=$?"; tail -n 1 gpurun_out/tests_wg2.log | cut -c1-160
DRB_PROFILE_DUMP=1 timeout 300 python bench.py --stage train --batch 1 --streams 1 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/train_dump_bf16_wg_model.txt; echo "dump new rc=$?"
DRB_WGRAD_SPLIT_MODEL=0 DRB_PROFILE_DUMP=1 timeout 300 python bench.py --stage train --batch 1 --streams 1 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/train_dump_bf16_wg_rule.txt; echo "dump old rc=$?"
timeout 600 python bench.py --stage train --batch 32 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/bench_train_wg_model.json 2> /dev/null; echo "train rc=$?"
python - <<PY
import json,re
def tot(p):
    w=0
    for l in open(p):
        m=re.search(r'M=\s*(-\d+) .*?: ([\d.]+) us',l)
        if m: w+=float(m[2])
    return w
print('wgrad per pair us: model', round(tot('gpurun_out/train_dump_bf16_wg_model.txt')), 'rule', round(tot('gpurun_out/train_dump_bf16_wg_rule.txt')))
d=json.loads(open('gpurun_out/bench_train_wg_model.json').read().strip().splitlines()[-1])
print('train', round(d['value'],2), d['unit'], 'e2e', round(d['e2e']['value'],2), 'ms/step', round(d['ms_per_step'],2))
PY
