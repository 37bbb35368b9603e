#!/bin/bash
mkdir -p gpurun_out
P="python -m pytest -q -p no:cacheprovider"
timeout 900 $P tests/test_extract_gpu.py -s > gpurun_out/tests_extract.log 2>&1; echo "extract tests rc=$?"; tail -n 3 gpurun_out/tests_extract.log; grep -E "128\^3 block|density mask:" gpurun_out/tests_extract.log | head
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_full_r02.json 2> gpurun_out/bench_full_r02.err; echo "full rc=$?"
DRB_MARCH_CELLMAJOR=0 timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_full_nocm_r02.json 2> gpurun_out/bench_full_nocm.err; echo "full (plain table) rc=$?"
timeout 900 python bench.py --stage train --batch 32 --steps 3 --warmup 3 > gpurun_out/bench_train_b32_r02.json 2> gpurun_out/bench_train.err; echo "train rc=$?"
timeout 600 python bench.py --stage batch --pairs-per-gpu 32 --steps 3 --warmup 3 --attention tc > gpurun_out/bench_batch32_r02.json 2> gpurun_out/bench_batch.err; echo "batch rc=$?"
timeout 900 python bench.py --stage batch --res 256 --max-tokens 20000 --attention tc --pairs-per-gpu 2 --steps 2 --warmup 3 > gpurun_out/bench_config5_r02.json 2> gpurun_out/bench_config5.err; echo "config5 rc=$?"; tail -n 3 gpurun_out/bench_config5.err
timeout 600 python bench.py --stage register --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_register_r02.json 2> gpurun_out/bench_register.err; echo "register rc=$?"
for f in full full_nocm train_b32 batch32 config5 register; do python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_${f}_r02.json').read().strip().splitlines()[-1])
    print('${f}', round(d['value'],2), d['unit'], 'e2e', round(d['e2e']['value'],2), 'ms/step', round(d['ms_per_step'],2), 'roofline frac', round(d['roofline']['frac'],4), d['config'].get('tokens'), d['roofline'].get('kernel_ms_per_launch'))
except Exception as e:
    print('${f}', 'ERR', e)
PY
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 5000 --csv --log-file gpurun_out/launches_train_r02.csv python scripts/time_train.py 128 bf16 1 > gpurun_out/ncu_train.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name regex:"wgrad_kernel|att_fwd_kernel|bn_bwd" -c 8 -o /tmp/r02_targets python scripts/ncu_targets.py > gpurun_out/ncu_targets.log 2>&1; echo "ncu full rc=$?"
ncu -i /tmp/r02_targets.ncu-rep --page raw --csv > gpurun_out/r02_targets_raw.csv 2>/dev/null; ls -la /tmp/r02_targets.ncu-rep gpurun_out/r02_targets_raw.csv
python scripts/ncu_keys.py < gpurun_out/r02_targets_raw.csv > gpurun_out/r02_ncu_targets_keys.txt 2>&1; wc -l gpurun_out/r02_ncu_targets_keys.txt
du -sh gpurun_out
