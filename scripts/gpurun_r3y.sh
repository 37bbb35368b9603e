#!/bin/bash
# final validation of the round-2 tree after the weight-gradient change (no ncu pass: 12 k launches cost 8 GPU-minutes)
mkdir -p gpurun_out
P="python -m pytest -q -p no:cacheprovider"
timeout 1800 $P tests -m gpu -x > gpurun_out/tests_gpu_all.log 2>&1; echo "gpu tests rc=$?"; tail -n 1 gpurun_out/tests_gpu_all.log | cut -c1-120
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "default bench rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 --streams 1 --no-cpu-baseline > gpurun_out/bench_full_s1.json 2> /dev/null; echo "full streams=1 rc=$?"
timeout 600 python bench.py --stage register --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_register.json 2> /dev/null; echo "register rc=$?"
timeout 600 python bench.py --stage batch --pairs-per-gpu 32 --steps 3 --warmup 3 > gpurun_out/bench_batch32_s4.json 2> /dev/null; echo "batch rc=$?"
timeout 900 python bench.py --stage train --batch 32 --steps 2 --warmup 3 > gpurun_out/bench_train_b32_s4.json 2> /dev/null; echo "train rc=$?"
timeout 900 python bench.py --stage batch --res 256 --max-tokens 20000 --pairs-per-gpu 4 --steps 2 --warmup 3 --streams 2 > gpurun_out/bench_config5_s2.json 2> /dev/null; echo "config5 rc=$?"
for f in default full_s1 register batch32_s4 train_b32_s4 config5_s2; do python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_${f}.json').read().strip().splitlines()[-1])
    print('${f}', round(d['value'],2), d['unit'], 'e2e', round(d['e2e']['value'],2), 'ms/step', round(d['ms_per_step'],2), 'roofline', round(d['roofline']['frac'],4), (d.get('roofline_register') or {}).get('tensor_pipe_frac_est'))
except Exception as e:
    print('${f}', 'ERR', e)
PY
done
timeout 240 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize_new_kernels.py wgrad > gpurun_out/r02_memcheck_wgrad.log 2>&1; echo "memcheck rc=$?"; grep -m3 "ERROR SUMMARY\|flag" gpurun_out/r02_memcheck_wgrad.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:wgrad -c 5 -f -o gpurun_out/r02_wgrad_staged python scripts/ncu_targets.py > gpurun_out/ncu_wgrad.log 2>&1; echo "ncu wgrad rc=$?"
ncu -i gpurun_out/r02_wgrad_staged.ncu-rep --page raw --csv 2>/dev/null | python scripts/ncu_keys.py > gpurun_out/r02_ncu_wgrad_staged_keys.txt; grep -c "Kernel Name" gpurun_out/r02_ncu_wgrad_staged_keys.txt
