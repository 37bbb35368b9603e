#!/bin/bash
mkdir -p gpurun_out
P="python -m pytest -q -p no:cacheprovider"
timeout 600 $P tests/test_pipeline_gpu.py -s > gpurun_out/tests_pipeline.log 2>&1; echo "pipeline tests rc=$?"; grep -E "passed|failed|worst" gpurun_out/tests_pipeline.log
for S in 4 8 12 16; do
  timeout 600 python bench.py --stage batch --pairs-per-gpu 32 --steps 3 --warmup 3 --attention tc --streams $S > gpurun_out/bench_batch32_s$S.json 2> gpurun_out/bench_batch32_s$S.err; echo "batch streams=$S rc=$?"; grep -v Warn gpurun_out/bench_batch32_s$S.err | tail -n 2
done
for S in 2 4; do
  timeout 900 python bench.py --stage train --batch 32 --steps 2 --warmup 3 --no-cpu-baseline --streams $S > gpurun_out/bench_train_b32_s$S.json 2> gpurun_out/bench_train_b32_s$S.err; echo "train streams=$S rc=$?"; grep -v Warn gpurun_out/bench_train_b32_s$S.err | tail -n 2
done
for f in batch32_s4 batch32_s8 batch32_s12 batch32_s16 train_b32_s2 train_b32_s4; do python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_${f}.json').read().strip().splitlines()[-1])
    print('${f}', round(d['value'],2), d['unit'], 'e2e', round(d['e2e']['value'],2), 'ms/step', round(d['ms_per_step'],2), 'roofline frac', round(d['roofline']['frac'],4))
except Exception as e:
    print('${f}', 'ERR', e)
PY
done
