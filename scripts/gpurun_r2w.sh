#!/bin/bash
mkdir -p gpurun_out
P="python -m pytest -q -p no:cacheprovider"
timeout 900 $P tests/test_edge_cases_gpu.py -q > gpurun_out/tests_edge.log 2>&1; echo "edge tests rc=$?"; tail -n 60 gpurun_out/tests_edge.log | grep -v "^$" | tail -45
