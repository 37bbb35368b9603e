#!/bin/bash
mkdir -p gpurun_out
nproc; free -g | head -2
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513"
timeout 600 $T bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_default_n8.json 2> gpurun_out/bench_default_n8.err; echo "default n8 rc=$?"
timeout 600 $T bench.py --gpus 8 --stage batch --pairs-per-gpu 32 --steps 3 --warmup 3 --attention tc > gpurun_out/bench_batch_n8.json 2> gpurun_out/bench_batch_n8.err; echo "batch n8 rc=$?"
for f in default_n8 batch_n8; do python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_${f}.json').read().strip().splitlines() if l.startswith('{')][-1])
    print('${f}', round(d['value'],2), d['unit'], 'n_gpus', d['n_gpus'], 'e2e', round(d['e2e']['value'],2), 'ms/step', round(d['ms_per_step'],2))
except Exception as e:
    print('${f}', 'ERR', e)
PY
done
grep -v Warn gpurun_out/bench_default_n8.err | grep -v "^\*\|OMP_NUM\|^$" | tail -5
