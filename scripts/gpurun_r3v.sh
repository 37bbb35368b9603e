#!/bin/bash
mkdir -p gpurun_out
P="python -m pytest -q -p no:cacheprovider"
timeout 1200 $P tests/test_kernels_gpu.py tests/test_forward_gpu.py tests/test_backward_gpu.py -q -x -k "igemm or forward_32 or forward_64 or forward_128 or stage_taps or backward_32 or backward_64 or bf16" > gpurun_out/tests_ws.log 2>&1; echo "tests rc=$?"; grep -E "^(FAILED|ERROR)" gpurun_out/tests_ws.log | head
for W in 1 0; do
DRB_IGEMM_WIDE_SPLIT=$W timeout 600 python bench.py --stage register --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_register_x.json 2> /dev/null
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_register_x.json').read().strip().splitlines()[-1])
print('wide_split=$W register fp32', round(d['value'],2), 'ms/step', round(d['ms_per_step'],3), 'igemm ms', round(d['roofline']['kernel_ms_per_step'],3))
PY
done
DRB_PROFILE_DUMP=1 timeout 600 python bench.py --stage register --precision bf16 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/igemm_dump_bf16_ws.txt
grep "M=8192 Cin=256 Cout=256 k=3\|M=1024 Cin=256 Cout=256 k=3\|M=8192 Cin=128 Cout=128 k=3" gpurun_out/igemm_dump_bf16_ws.txt | head -4
timeout 600 python bench.py --stage batch --pairs-per-gpu 32 --steps 3 --warmup 3 > gpurun_out/bench_batch32_s4.json 2> /dev/null
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_batch32_s4.json').read().strip().splitlines()[-1])
print('batch', round(d['value'],2), 'e2e', round(d['e2e']['value'],2))
PY
