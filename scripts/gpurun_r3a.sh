#!/bin/bash
mkdir -p gpurun_out
P="python -m pytest -q -p no:cacheprovider"
timeout 120 python scripts/att_once.py 1 300; timeout 120 python scripts/att_once.py 1 1000
timeout 600 $P tests/test_attention_gpu.py -q -s > gpurun_out/tests_att.log 2>&1; echo "attention tests rc=$?"; grep -E "planes=1|passed|failed" gpurun_out/tests_att.log | tail -8
timeout 300 python scripts/att_time.py 8192
DRB_ATT_TWO_TILES=0 timeout 300 python scripts/att_time.py 8192 | head -1
