#!/bin/bash
mkdir -p gpurun_out
P="python -m pytest -q -p no:cacheprovider"
timeout 600 $P tests/test_attention_gpu.py -s > gpurun_out/att.log 2>&1; echo "attention rc=$?"; grep -E "tcgen05 att|forward with|passed|failed|watchdog|sites" gpurun_out/att.log | head -30
timeout 1800 $P tests -m gpu -s --deselect tests/test_attention_gpu.py > gpurun_out/gpu_all.log 2>&1
echo "gpu tests rc=$?"; tail -n 4 gpurun_out/gpu_all.log
grep -E "SMALLEST|pair, tensor|digests checked|tensors within" gpurun_out/gpu_all.log | head
timeout 600 python bench.py --stage register --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_register.json 2> gpurun_out/bench_register.err; echo "register rc=$?"; head -c 300 gpurun_out/bench_register.json; echo
DRB_TC_ATTENTION=1 timeout 600 python bench.py --stage register --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_register_tc.json 2> gpurun_out/bench_register_tc.err; echo "register tc rc=$?"; head -c 300 gpurun_out/bench_register_tc.json; echo
