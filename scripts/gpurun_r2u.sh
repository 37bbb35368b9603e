#!/bin/bash
mkdir -p gpurun_out
P="python -m pytest -q -p no:cacheprovider"
timeout 1800 $P tests -m gpu -x > gpurun_out/tests_gpu_all.log 2>&1; echo "gpu tests rc=$?"; tail -n 3 gpurun_out/tests_gpu_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "default bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_default.json').read().strip().splitlines()[-1])
print('default', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'ms/step', round(d['ms_per_step'],2), d['config'].get('streams'), 'roofline', round(d['roofline']['frac'],5), d['roofline']['kernel_ms_per_launch'], 'cpu', d.get('cpu_baseline',{}).get('value'))
r=json.loads(open('gpurun_out/bench_reference.json').read().strip().splitlines()[-1])
print('reference', r['value'], r.get('native_so_loaded'))
PY
