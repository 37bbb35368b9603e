#!/bin/bash
# evidence: ncu --set full of the implicit-GEMM kernels at five shapes + fresh per-shape tables of the final tree
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:igemm -c 5 -f -o gpurun_out/r02_igemm_final python scripts/ncu_igemm_targets.py > gpurun_out/ncu_igemm.log 2>&1; echo "ncu igemm rc=$?"; tail -n 7 gpurun_out/ncu_igemm.log
ncu -i gpurun_out/r02_igemm_final.ncu-rep --page raw --csv 2>/dev/null | python scripts/ncu_keys.py > gpurun_out/r02_ncu_igemm_keys.txt; grep -c "Kernel Name" gpurun_out/r02_ncu_igemm_keys.txt
DRB_PROFILE_DUMP=1 timeout 300 python bench.py --stage register --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/igemm_dump_fp32_final.txt; echo "rc=$?"
DRB_PROFILE_DUMP=1 timeout 300 python bench.py --stage register --precision bf16 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/igemm_dump_bf16_final.txt; echo "rc=$?"
python scripts/igemm_per_shape.py gpurun_out/igemm_dump_fp32_final.txt:3 gpurun_out/igemm_dump_bf16_final.txt:1 > gpurun_out/r02_igemm_per_shape_final.txt; head -3 gpurun_out/r02_igemm_per_shape_final.txt
