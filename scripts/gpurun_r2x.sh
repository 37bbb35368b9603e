#!/bin/bash
mkdir -p gpurun_out
P="python -m pytest -q -p no:cacheprovider"
timeout 900 $P tests/test_forward_gpu.py -q -s -k "bf16" > gpurun_out/tests_bf16.log 2>&1; echo "bf16 tests rc=$?"; grep -E "^.?F?bf16 |passed|failed|AssertionError" gpurun_out/tests_bf16.log | tail
