#!/bin/bash
mkdir -p gpurun_out
P="python -m pytest -q -p no:cacheprovider"
timeout 1500 $P tests -m gpu -s > gpurun_out/gpu_all.log 2>&1
echo "gpu tests rc=$?"; tail -n 12 gpurun_out/gpu_all.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 4 gpurun_out/smoke.log
timeout 600 python scripts/time_train.py 128 bf16 5 > gpurun_out/time_train_bf16.log 2>&1; tail -n 3 gpurun_out/time_train_bf16.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "full rc=$?"
head -c 700 gpurun_out/bench_full.json; echo
