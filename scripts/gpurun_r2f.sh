#!/bin/bash
mkdir -p gpurun_out
P="python -m pytest -q -p no:cacheprovider"
timeout 600 $P tests/test_attention_gpu.py -s -x > gpurun_out/att.log 2>&1; echo "attention rc=$?"; grep -E "tcgen05|forward with|passed|failed|watchdog|sites" gpurun_out/att.log | head -30
timeout 900 $P tests/test_backward_gpu.py -s -k "backward_32" > gpurun_out/bwd32.log 2>&1; echo "bwd32 rc=$?"; grep -E "MEDIAN|pair, tensor|digests|passed|failed" gpurun_out/bwd32.log | head
timeout 900 python bench.py --stage train --batch 32 --steps 3 --warmup 3 > gpurun_out/bench_train_b32.json 2> gpurun_out/bench_train_b32.err; echo "train rc=$?"
timeout 600 python bench.py --stage batch --pairs-per-gpu 32 --steps 3 --warmup 3 > gpurun_out/bench_batch32.json 2> gpurun_out/bench_batch32.err; echo "batch rc=$?"
head -c 400 gpurun_out/bench_train_b32.json; echo; head -c 400 gpurun_out/bench_batch32.json; echo
tail -n 3 gpurun_out/bench_train_b32.err gpurun_out/bench_batch32.err
