#!/bin/bash
# stream pipeline: parity tests, then batch / train stages at several stream counts
mkdir -p gpurun_out
P="python -m pytest -q -p no:cacheprovider"
timeout 600 $P tests/test_pipeline_gpu.py -s > gpurun_out/tests_pipeline.log 2>&1; echo "pipeline tests rc=$?"; tail -n 15 gpurun_out/tests_pipeline.log
for S in 1 2 4 8; do
  timeout 600 python bench.py --stage batch --pairs-per-gpu 32 --steps 3 --warmup 3 --attention tc --streams $S > gpurun_out/bench_batch32_s$S.json 2> gpurun_out/bench_batch32_s$S.err; echo "batch streams=$S rc=$?"; tail -n 2 gpurun_out/bench_batch32_s$S.err
done
for S in 1 4; do
  timeout 900 python bench.py --stage train --batch 32 --steps 2 --warmup 3 --no-cpu-baseline --streams $S > gpurun_out/bench_train_b32_s$S.json 2> gpurun_out/bench_train_b32_s$S.err; echo "train streams=$S rc=$?"; tail -n 2 gpurun_out/bench_train_b32_s$S.err
done
for f in batch32_s1 batch32_s2 batch32_s4 batch32_s8 train_b32_s1 train_b32_s4; do python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_${f}.json').read().strip().splitlines()[-1])
    print('${f}', round(d['value'],2), d['unit'], 'e2e', round(d['e2e']['value'],2), 'ms/step', round(d['ms_per_step'],2), 'roofline frac', round(d['roofline']['frac'],4))
except Exception as e:
    print('${f}', 'ERR', e)
PY
done
nvidia-smi --query-gpu=memory.used --format=csv
