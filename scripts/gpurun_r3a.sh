#!/bin/bash
mkdir -p gpurun_out
P="python -m pytest -q -p no:cacheprovider"
timeout 600 $P tests/test_attention_gpu.py -q > gpurun_out/tests_att.log 2>&1; echo "attention tests rc=$?"; tail -n 2 gpurun_out/tests_att.log
timeout 300 python scripts/att_time.py 8192
