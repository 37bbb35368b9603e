#!/bin/bash
mkdir -p gpurun_out
P="python -m pytest -q -p no:cacheprovider"
timeout 900 $P tests/test_attention_gpu.py tests/test_forward_gpu.py tests/test_backward_gpu.py -s > gpurun_out/tests_sub.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/tests_sub.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_full_r02.json 2> gpurun_out/bench_full_r02.err; echo "full rc=$?"
timeout 900 python bench.py --stage train --batch 32 --steps 3 --warmup 3 > gpurun_out/bench_train_b32_r02.json 2> gpurun_out/bench_train.err; echo "train rc=$?"
timeout 600 python bench.py --stage batch --pairs-per-gpu 32 --steps 3 --warmup 3 --attention tc > gpurun_out/bench_batch32_r02.json 2> gpurun_out/bench_batch.err; echo "batch rc=$?"
timeout 900 python bench.py --stage batch --res 256 --max-tokens 16384 --attention tc --pairs-per-gpu 2 --steps 2 --warmup 3 > gpurun_out/bench_config5_r02.json 2> gpurun_out/bench_config5.err; echo "config5 rc=$?"; tail -n 3 gpurun_out/bench_config5.err
timeout 600 python bench.py --stage register --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_register_r02.json 2> gpurun_out/bench_register.err; echo "register rc=$?"
for f in full train_b32 batch32 config5 register; do python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_${f}_r02.json').read().strip().splitlines()[-1])
    print('${f}', round(d['value'],2), d['unit'], 'e2e', round(d['e2e']['value'],2), 'ms/step', round(d['ms_per_step'],2), 'roofline frac', round(d['roofline']['frac'],4), d['config'].get('tokens'))
except Exception as e:
    print('${f}', 'ERR', e)
PY
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 5000 --csv --log-file gpurun_out/launches_train_r02.csv python scripts/time_train.py 128 bf16 1 > gpurun_out/ncu_train.log 2>&1; echo "ncu list rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name regex:"wgrad_kernel|att_fwd_kernel" --launch-skip 200 --launch-count 110 -o gpurun_out/r02_wgrad_att python scripts/time_train.py 128 bf16 1 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"; ls -la gpurun_out/*.ncu-rep
