#!/bin/bash
mkdir -p gpurun_out
P="python -m pytest -q -p no:cacheprovider"
timeout 900 $P tests/test_kernels_gpu.py tests/test_forward_gpu.py -q -x -k "maxpool or forward_64 or forward_128 or stage_taps or forward_32" > gpurun_out/tests_mp.log 2>&1; echo "tests rc=$?"; grep -E "^(FAILED|ERROR)" gpurun_out/tests_mp.log | head
timeout 600 python bench.py --stage register --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_register.json 2> /dev/null; echo "register rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_register.json').read().strip().splitlines()[-1])
print('register', round(d['value'],2), 'ms/step', round(d['ms_per_step'],3))
PY
