#!/bin/bash
mkdir -p gpurun_out
P="python -m pytest -q -p no:cacheprovider"
timeout 900 $P tests/test_kernels_gpu.py -q -x -s -k "two_tile" > gpurun_out/tests_ig2.log 2>&1; echo "two-tile test rc=$?"; grep -E "Error|error|assert|passed|failed" gpurun_out/tests_ig2.log | head -8
