#!/bin/bash
# two-GPU sanity of the final tree: default (full path) and the sharded bf16 batch, launched as the driver does
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $T bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_default_n2_final.json 2> gpurun_out/bench_default_n2_final.err; echo "default n2 rc=$?"
timeout 600 $T bench.py --gpus 2 --stage batch --pairs-per-gpu 16 --steps 2 --warmup 3 > gpurun_out/bench_batch_n2_final.json 2> gpurun_out/bench_batch_n2_final.err; echo "batch n2 rc=$?"
for f in default_n2_final batch_n2_final; do python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_${f}.json').read().strip().splitlines()[-1])
    print('${f}', round(d['value'],2), d['unit'], 'n_gpus', d['n_gpus'], 'e2e', round(d['e2e']['value'],2), 'ms/step', round(d['ms_per_step'],2))
except Exception as e:
    print('${f}', 'ERR', e)
PY
done
tail -n 3 gpurun_out/bench_default_n2_final.err
