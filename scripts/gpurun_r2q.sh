#!/bin/bash
mkdir -p gpurun_out
P="python -m pytest -q -p no:cacheprovider"
timeout 900 $P tests/test_kernels_gpu.py -x -q -k "batchnorm" > gpurun_out/tests_bn.log 2>&1; echo "bn tests rc=$?"; tail -n 3 gpurun_out/tests_bn.log
timeout 900 $P tests/test_extract_gpu.py tests/test_forward_gpu.py tests/test_backward_gpu.py -x -q > gpurun_out/tests_efb.log 2>&1; echo "extract+forward+backward tests rc=$?"; tail -n 3 gpurun_out/tests_efb.log
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_full_r02c.json 2> gpurun_out/bench_full_r02c.err; echo "full rc=$?"
DRB_BN_SMALL=0 timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_full_r02c_nobnsmall.json 2> /dev/null; echo "full (general BN) rc=$?"
timeout 900 python bench.py --stage register --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_register_r02c.json 2> /dev/null; echo "register rc=$?"
for f in full_r02c full_r02c_nobnsmall register_r02c; do python - <<PY
import json
d=json.loads(open('gpurun_out/bench_${f}.json').read().strip().splitlines()[-1])
print('${f}', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'ms/step', round(d['ms_per_step'],2), (d['config'].get('extract') or {}).get('stage_ms'), d['gpu_launches'])
PY
done
