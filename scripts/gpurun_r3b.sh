#!/bin/bash
mkdir -p gpurun_out
P="python -m pytest -q -p no:cacheprovider"
timeout 900 $P tests/test_attention_gpu.py tests/test_forward_gpu.py tests/test_pipeline_gpu.py -q > gpurun_out/tests_afp.log 2>&1; echo "attention+forward+pipeline tests rc=$?"; tail -n 2 gpurun_out/tests_afp.log
timeout 900 python bench.py --stage batch --res 256 --max-tokens 20000 --pairs-per-gpu 4 --steps 2 --warmup 3 --streams 2 > gpurun_out/bench_config5_s2.json 2> gpurun_out/bench_config5_s2.err; echo "config5 rc=$?"
timeout 600 python bench.py --stage batch --pairs-per-gpu 32 --steps 3 --warmup 3 > gpurun_out/bench_batch32_s4.json 2> /dev/null; echo "batch rc=$?"
for f in config5_s2 batch32_s4; do python - <<PY
import json
d=json.loads(open('gpurun_out/bench_${f}.json').read().strip().splitlines()[-1])
print('${f}', round(d['value'],2), d['unit'], 'e2e', round(d['e2e']['value'],2), 'ms/step', round(d['ms_per_step'],2), d['config'].get('tokens'))
PY
done
