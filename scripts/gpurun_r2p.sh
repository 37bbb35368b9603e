#!/bin/bash
# two GPUs of one box: the three bench stages through torchrun (NCCL), as the driver launches them
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $T bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_full_n2.json 2> gpurun_out/bench_full_n2.err; echo "full n2 rc=$?"
timeout 600 $T bench.py --gpus 2 --stage batch --pairs-per-gpu 32 --steps 3 --warmup 3 --attention tc > gpurun_out/bench_batch_n2.json 2> gpurun_out/bench_batch_n2.err; echo "batch n2 rc=$?"
timeout 600 $T bench.py --gpus 2 --stage batch --total-pairs 64 --steps 3 --warmup 3 --attention tc > gpurun_out/bench_batch_strong_n2.json 2> gpurun_out/bench_batch_strong_n2.err; echo "batch strong n2 rc=$?"
timeout 900 $T bench.py --gpus 2 --stage train --batch 16 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_train_n2.json 2> gpurun_out/bench_train_n2.err; echo "train n2 rc=$?"
timeout 300 $T bench.py --gpus 2 --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; echo "reference n2 rc=$?"
for f in full_n2 batch_n2 batch_strong_n2 train_n2; do python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_${f}.json').read().strip().splitlines() if l.startswith('{')][-1])
    print('${f}', round(d['value'],2), d['unit'], 'n_gpus', d['n_gpus'], 'e2e', round(d['e2e']['value'],2), 'ms/step', round(d['ms_per_step'],2), d['scaling'])
except Exception as e:
    print('${f}', 'ERR', e)
PY
done
tail -c 600 gpurun_out/bench_ref_n2.json
