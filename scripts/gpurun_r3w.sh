#!/bin/bash
# weight gradient: staged vector reductions + taps side by side in the N tile - tests, A/B on the training step, per-shape dump
mkdir -p gpurun_out
P="python -m pytest -q -p no:cacheprovider"
timeout 900 $P tests/test_backward_gpu.py tests/test_pipeline_gpu.py tests/test_abi.py -m gpu -x -s > gpurun_out/tests_wg.log 2>&1; echo "tests rc=$?"; tail -n 2 gpurun_out/tests_wg.log | cut -c1-160
grep "wgrad" gpurun_out/tests_wg.log | cut -c1-170 | tail -n 40
timeout 600 python bench.py --stage train --batch 32 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_train_wg_new.json 2> /dev/null; echo "train new rc=$?"
DRB_WGRAD_STAGE=0 DRB_WGRAD_FLAT_N=0 timeout 600 python bench.py --stage train --batch 32 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_train_wg_old.json 2> /dev/null; echo "train old rc=$?"
DRB_WGRAD_FLAT_N=0 timeout 600 python bench.py --stage train --batch 32 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_train_wg_stage_only.json 2> /dev/null; echo "train stage-only rc=$?"
DRB_PROFILE_DUMP=1 timeout 600 python bench.py --stage train --batch 1 --streams 1 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/train_dump_bf16_wg.txt; echo "dump rc=$?"
for f in new old stage_only; do python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_train_wg_${f}.json').read().strip().splitlines()[-1])
    print('${f}', round(d['value'],2), d['unit'], 'e2e', round(d['e2e']['value'],2), 'ms/step', round(d['ms_per_step'],2))
except Exception as e:
    print('${f}', 'ERR', e)
PY
done
