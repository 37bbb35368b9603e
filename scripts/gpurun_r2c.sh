#!/bin/bash
mkdir -p gpurun_out
P="python -m pytest -q -p no:cacheprovider"
timeout 1500 $P tests -m gpu -s > gpurun_out/gpu_all.log 2>&1
echo "gpu tests rc=$?"; tail -n 3 gpurun_out/gpu_all.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 4 gpurun_out/smoke.log
timeout 900 python bench.py --stage train --batch 4 --steps 3 --warmup 3 > gpurun_out/bench_train_b4.json 2> gpurun_out/bench_train_b4.err; echo "train rc=$?"
timeout 600 python bench.py --stage batch --pairs-per-gpu 8 --steps 3 --warmup 3 > gpurun_out/bench_batch8.json 2> gpurun_out/bench_batch8.err; echo "batch rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "full rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_train.csv python scripts/time_train.py 128 bf16 1 > gpurun_out/ncu_train.log 2>&1; echo "ncu rc=$?"
head -c 1500 gpurun_out/bench_train_b4.json; echo; head -c 600 gpurun_out/bench_batch8.json; echo; head -c 400 gpurun_out/bench_full.json; echo
