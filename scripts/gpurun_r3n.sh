#!/bin/bash
mkdir -p gpurun_out
timeout 500 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize_new_kernels.py all > gpurun_out/r02_memcheck_new_kernels.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|Invalid|done" gpurun_out/r02_memcheck_new_kernels.log | head -5
timeout 500 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 20 python scripts/sanitize_new_kernels.py att > gpurun_out/r02_racecheck_attention.log 2>&1; echo "racecheck att rc=$?"; grep -E "RACECHECK SUMMARY|hazard|done" gpurun_out/r02_racecheck_attention.log | head -5
timeout 500 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 20 python scripts/sanitize_new_kernels.py elem > gpurun_out/r02_racecheck_elementwise.log 2>&1; echo "racecheck elem rc=$?"; grep -E "RACECHECK SUMMARY|hazard|done" gpurun_out/r02_racecheck_elementwise.log | head -5
