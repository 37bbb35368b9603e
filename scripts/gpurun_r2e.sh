#!/bin/bash
mkdir -p gpurun_out
P="python -m pytest -q -p no:cacheprovider"
timeout 600 $P tests/test_attention_gpu.py -s > gpurun_out/att.log 2>&1; echo "attention rc=$?"; grep -E "tcgen05|forward with|passed|failed|Error" gpurun_out/att.log | head -30
timeout 1800 $P tests -m gpu -s --deselect tests/test_attention_gpu.py > gpurun_out/gpu_all.log 2>&1
echo "gpu tests rc=$?"; tail -n 6 gpurun_out/gpu_all.log
grep -E "within|mismatch|MEDIAN|pair, tensor|digests checked" gpurun_out/gpu_all.log | head -30
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 3 gpurun_out/smoke.log
