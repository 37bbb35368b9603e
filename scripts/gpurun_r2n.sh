#!/bin/bash
# round-2 evidence: the whole GPU test suite, smoke, bench lines, the ncu launch list of the full path and
# `ncu --set full` captures of the marcher (cell-major) and of the rewritten attention kernel
mkdir -p gpurun_out
P="python -m pytest -q -p no:cacheprovider"
timeout 1500 $P tests -m gpu -x > gpurun_out/tests_gpu_all.log 2>&1; echo "gpu tests rc=$?"; tail -n 3 gpurun_out/tests_gpu_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_full_r02b.json 2> gpurun_out/bench_full_r02b.err; echo "full rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_full_r02.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu list rc=$?"
SEED=501 timeout 600 ncu --set full --clock-control none --import-source on -k regex:surface_mask -s 2 -c 1 -f -o /tmp/surface_full python scripts/extract_once.py 3 501 > gpurun_out/ncu_surface.log 2>&1; echo "ncu surface rc=$?"
ncu -i /tmp/surface_full.ncu-rep --page raw --csv 2>/dev/null | python scripts/ncu_keys.py > gpurun_out/r02_ncu_surface_seed501_keys.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:att_fwd_kernel -s 2 -c 2 -f -o /tmp/att_full python scripts/att_time.py 8192 > gpurun_out/ncu_att.log 2>&1; echo "ncu att rc=$?"
ncu -i /tmp/att_full.ncu-rep --page raw --csv 2>/dev/null | python scripts/ncu_keys.py > gpurun_out/r02_ncu_attention_keys.txt
wc -l gpurun_out/r02_ncu_*keys.txt
du -sh gpurun_out
