#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:att_fwd2_kernel -s 2 -c 1 -f -o /tmp/att2_full python scripts/att_time.py 8192 > gpurun_out/ncu_att2.log 2>&1; echo "ncu att2 rc=$?"
ncu -i /tmp/att2_full.ncu-rep --page raw --csv 2>/dev/null | python scripts/ncu_keys.py > gpurun_out/r02_ncu_attention2_keys.txt; wc -l gpurun_out/r02_ncu_attention2_keys.txt
