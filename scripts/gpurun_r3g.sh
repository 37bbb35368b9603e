#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_register_bf16.csv python bench.py --stage register --precision bf16 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_reg_bf16.log 2>&1; echo "ncu list rc=$?"
