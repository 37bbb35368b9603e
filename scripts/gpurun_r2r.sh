#!/bin/bash
mkdir -p gpurun_out
DRB_PROFILE_DUMP=1 timeout 600 python bench.py --stage register --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/igemm_dump_fp32.txt; echo "rc=$?"
DRB_PROFILE_DUMP=1 timeout 600 python bench.py --stage register --precision bf16 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/igemm_dump_bf16.txt; echo "rc=$?"
grep -c igemm gpurun_out/igemm_dump_fp32.txt gpurun_out/igemm_dump_bf16.txt
