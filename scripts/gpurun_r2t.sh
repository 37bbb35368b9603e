#!/bin/bash
mkdir -p gpurun_out
P="python -m pytest -q -p no:cacheprovider"
timeout 900 $P tests/test_pipeline_gpu.py -x -q -s > gpurun_out/tests_pipeline2.log 2>&1; echo "pipeline tests rc=$?"; grep -E "passed|failed|masked|Error|error" gpurun_out/tests_pipeline2.log | tail -8
