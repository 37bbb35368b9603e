#!/bin/bash
mkdir -p gpurun_out
P="python -m pytest -q -p no:cacheprovider"
timeout 1200 $P tests/test_kernels_gpu.py tests/test_forward_gpu.py tests/test_backward_gpu.py -q -x -k "im2col or batchnorm or forward_64 or forward_128 or stage_taps or backward_32 or backward_64 or training_steps" > gpurun_out/tests_i2c.log 2>&1; echo "tests rc=$?"; grep -E "^(FAILED|ERROR)" gpurun_out/tests_i2c.log | head -5
timeout 600 python bench.py --stage register --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_register_r02f.json 2> /dev/null; echo "register rc=$?"
timeout 600 python bench.py --stage batch --pairs-per-gpu 32 --steps 3 --warmup 3 > gpurun_out/bench_batch32_s4.json 2> /dev/null; echo "batch rc=$?"
timeout 900 python bench.py --stage train --batch 32 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_train_b32_s4.json 2> /dev/null; echo "train rc=$?"
for f in register_r02f batch32_s4 train_b32_s4; do python - <<PY
import json
d=json.loads(open('gpurun_out/bench_${f}.json').read().strip().splitlines()[-1])
print('${f}', round(d['value'],2), d['unit'], 'e2e', round(d['e2e']['value'],2), 'ms/step', round(d['ms_per_step'],2))
PY
done
