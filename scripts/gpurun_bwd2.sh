#!/bin/bash
mkdir -p gpurun_out
P="python -m pytest -q -s -p no:cacheprovider"
timeout 900 $P tests/test_backward_gpu.py -k "backward_32 or backward_64 or training_steps" > gpurun_out/bwd_full.log 2>&1
echo "full rc=$?"
timeout 600 python scripts/time_train.py 128 bf16 5 > gpurun_out/time_train_bf16.log 2>&1
echo "time bf16 rc=$?"
timeout 600 python scripts/time_train.py 128 fp32 3 > gpurun_out/time_train_fp32.log 2>&1
echo "time fp32 rc=$?"
tail -n 30 gpurun_out/bwd_full.log; tail -n 12 gpurun_out/time_train_bf16.log gpurun_out/time_train_fp32.log
