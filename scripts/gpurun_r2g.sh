#!/bin/bash
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python scripts/att_once.py 2 300 > gpurun_out/att_sanitizer.log 2>&1; echo "sanitizer rc=$?"
grep -v "^=========     Host Frame\|^=========         in \|^=========                in " gpurun_out/att_sanitizer.log | head -60
