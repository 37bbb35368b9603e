#!/bin/bash
# repeat of the GPU suite + smoke on a fresh box (flakiness check of the final tree)
mkdir -p gpurun_out
P="python -m pytest -q -p no:cacheprovider"
timeout 900 $P tests -m gpu -x > gpurun_out/tests_gpu_all_repeat.log 2>&1; echo "gpu tests rc=$?"; tail -n 1 gpurun_out/tests_gpu_all_repeat.log | cut -c1-120
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_repeat.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/smoke_repeat.log
