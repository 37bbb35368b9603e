#!/bin/bash
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
timeout 600 $T bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_default_n2.json 2> gpurun_out/bench_default_n2.err; echo "default n2 rc=$?"
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_default_n1.json 2> gpurun_out/bench_default_n1.err; echo "default n1 rc=$?"
for f in default_n1 default_n2; do python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_${f}.json').read().strip().splitlines() if l.startswith('{')][-1])
    print('${f}', round(d['value'],2), d['unit'], 'n_gpus', d['n_gpus'], 'e2e', round(d['e2e']['value'],2), 'ms/step', round(d['ms_per_step'],2), d['roofline']['kernel_ms_per_launch'])
except Exception as e:
    print('${f}', 'ERR', e)
PY
done
grep -v Warn gpurun_out/bench_default_n2.err | tail -5
