#!/bin/bash
mkdir -p gpurun_out
DRB_PROFILE_DUMP=1 timeout 600 python bench.py --stage train --batch 1 --streams 1 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/train_dump_bf16.txt; echo "rc=$?"
grep -c igemm gpurun_out/train_dump_bf16.txt
