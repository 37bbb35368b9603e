#!/bin/bash
# round 2: first GPU pass over the backward kernels (each group in its own process: a trapped kernel must not
# take the other groups down)
mkdir -p gpurun_out
P="python -m pytest -q -s -p no:cacheprovider"
timeout 900 $P tests/test_backward_gpu.py -k "not wgrad and not backward_32 and not backward_64 and not training_steps" > gpurun_out/bwd_elem.log 2>&1
echo "elem rc=$?"
timeout 600 $P tests/test_backward_gpu.py -k "wgrad" > gpurun_out/bwd_wgrad.log 2>&1
echo "wgrad rc=$?"
DRB_WGRAD_DESC_SWAP=1 timeout 600 $P tests/test_backward_gpu.py -k "wgrad" > gpurun_out/bwd_wgrad_swap.log 2>&1
echo "wgrad swap rc=$?"
timeout 900 $P tests/test_backward_gpu.py -k "backward_32 or backward_64 or training_steps" > gpurun_out/bwd_full.log 2>&1
echo "full rc=$?"
timeout 1500 $P tests -m gpu --ignore=tests/test_backward_gpu.py > gpurun_out/fwd_all.log 2>&1
echo "fwd rc=$?"
tail -n 5 gpurun_out/bwd_elem.log gpurun_out/bwd_wgrad.log gpurun_out/bwd_wgrad_swap.log gpurun_out/bwd_full.log gpurun_out/fwd_all.log
