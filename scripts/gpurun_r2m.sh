#!/bin/bash
mkdir -p gpurun_out
P="python -m pytest -q -p no:cacheprovider"
timeout 300 python scripts/att_once.py 1 300; timeout 300 python scripts/att_once.py 2 1000
timeout 600 $P tests/test_attention_gpu.py -s > gpurun_out/tests_att.log 2>&1; echo "attention tests rc=$?"; grep -E "passed|failed|rel err|tcgen05 att|forward with" gpurun_out/tests_att.log | tail -20
timeout 300 python scripts/att_time.py 8192
timeout 900 python bench.py --stage batch --res 256 --max-tokens 20000 --attention tc --pairs-per-gpu 4 --steps 2 --warmup 3 --streams 2 > gpurun_out/bench_config5_s2.json 2> gpurun_out/bench_config5_s2.err; echo "config5 rc=$?"; grep -v Warn gpurun_out/bench_config5_s2.err | tail -n 3
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_config5_s2.json').read().strip().splitlines()[-1])
print('config5', round(d['value'],2), d['unit'], 'e2e', round(d['e2e']['value'],2), 'ms/step', round(d['ms_per_step'],2), d['config'].get('tokens'))
PY
