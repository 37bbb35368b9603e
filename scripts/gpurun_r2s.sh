#!/bin/bash
mkdir -p gpurun_out
for S in 2 3 4; do
timeout 900 python bench.py --steps 24 --warmup 3 --no-cpu-baseline --streams $S > gpurun_out/bench_full_s$S.json 2> gpurun_out/bench_full_s$S.err; echo "full streams=$S rc=$?"; grep -v Warn gpurun_out/bench_full_s$S.err | tail -n 3
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_full_s$S.json').read().strip().splitlines()[-1])
print('full s$S', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'ms/step', round(d['ms_per_step'],2), d['config']['extract']['stage_ms'], d['roofline']['kernel_ms_per_launch'])
PY
done
