#!/bin/bash
# BASELINE configs[4] as written: 256^3 grids, ~8k-token clouds (19.4 k tokens per pair), batch 64 on 8 GPUs
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514"
OMP_NUM_THREADS=4 timeout 900 $T bench.py --gpus 8 --stage batch --res 256 --max-tokens 20000 --pairs-per-gpu 8 --streams 2 --steps 2 --warmup 3 > gpurun_out/bench_config5_n8.json 2> gpurun_out/bench_config5_n8.err; echo "config5 n8 rc=$?"
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_config5_n8.json').read().strip().splitlines() if l.startswith('{')][-1])
    print('config5_n8', round(d['value'],2), d['unit'], 'n_gpus', d['n_gpus'], 'e2e', round(d['e2e']['value'],2), 'ms/step', round(d['ms_per_step'],2), d['config'].get('tokens'), d['config']['workload'])
except Exception as e:
    print('ERR', e)
PY
grep -v Warn gpurun_out/bench_config5_n8.err | grep -v "^\*\|OMP_NUM\|^$" | tail -5
