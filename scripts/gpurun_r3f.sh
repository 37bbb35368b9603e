#!/bin/bash
mkdir -p gpurun_out
P="python -m pytest -q -p no:cacheprovider"
timeout 1200 $P tests/test_forward_gpu.py tests/test_backward_gpu.py tests/test_kernels_gpu.py -q -x > gpurun_out/tests_fbk.log 2>&1; echo "forward+backward+kernels tests rc=$?"; grep -E "^(FAILED|ERROR)|assert" gpurun_out/tests_fbk.log | head -5
timeout 600 python bench.py --stage register --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_register_r02e.json 2> /dev/null; echo "register rc=$?"
DRB_FUSE_TOPDOWN=0 timeout 600 python bench.py --stage register --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_register_r02e_nofuse.json 2> /dev/null; echo "register (separate merge) rc=$?"
timeout 600 python bench.py --stage batch --pairs-per-gpu 32 --steps 3 --warmup 3 > gpurun_out/bench_batch32_s4.json 2> /dev/null; echo "batch rc=$?"
for f in register_r02e register_r02e_nofuse batch32_s4; do python - <<PY
import json
d=json.loads(open('gpurun_out/bench_${f}.json').read().strip().splitlines()[-1])
print('${f}', round(d['value'],2), d['unit'], 'e2e', round(d['e2e']['value'],2), 'ms/step', round(d['ms_per_step'],2))
PY
done
