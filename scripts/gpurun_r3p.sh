#!/bin/bash
mkdir -p gpurun_out
P="python -m pytest -q -p no:cacheprovider"
timeout 1200 $P tests/test_forward_gpu.py tests/test_backward_gpu.py tests/test_pipeline_gpu.py -q -x -k "bf16 or training_steps or pipeline" > gpurun_out/tests_bf16b.log 2>&1; echo "bf16 tests rc=$?"; grep -E "^(FAILED|ERROR)" gpurun_out/tests_bf16b.log | head
DRB_PROFILE_DUMP=1 timeout 600 python bench.py --stage register --precision bf16 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/igemm_dump_bf16_two.txt; echo "rc=$?"
grep "M= *524288\|M= *65536 Cin=  256 Cout=  256 k=3" gpurun_out/igemm_dump_bf16_two.txt | head -6
timeout 600 python bench.py --stage batch --pairs-per-gpu 32 --steps 3 --warmup 3 > gpurun_out/bench_batch32_s4.json 2> /dev/null; echo "batch rc=$?"
timeout 900 python bench.py --stage train --batch 32 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_train_b32_s4.json 2> /dev/null; echo "train rc=$?"
timeout 900 python bench.py --stage batch --res 256 --max-tokens 20000 --pairs-per-gpu 4 --steps 2 --warmup 3 --streams 2 > gpurun_out/bench_config5_s2.json 2> /dev/null; echo "config5 rc=$?"
for f in batch32_s4 train_b32_s4 config5_s2; do python - <<PY
import json
d=json.loads(open('gpurun_out/bench_${f}.json').read().strip().splitlines()[-1])
print('${f}', round(d['value'],2), d['unit'], 'e2e', round(d['e2e']['value'],2), 'ms/step', round(d['ms_per_step'],2), 'roofline', round(d['roofline']['frac'],4))
PY
done
