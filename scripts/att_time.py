"""Device time of the tcgen05 attention kernel at the 8k-token size of BASELINE configs[4] (CUDA events, 20 calls)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dreg_nerf_b200 as pkg  # noqa: F401
from importlib import import_module
ops = import_module("dreg-nerf_b200.ops")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
torch.manual_seed(0)
qkv = torch.randn(2 * n, 768).cuda()
for planes in (1, 2):
    for _ in range(3):
        ops.mha_tc(qkv, n, [(0, 1)], planes=planes)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.mha_tc(qkv, n, [(0, 1)], planes=planes)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    fl = 4.0 * n * n * 32 * 8
    print("planes %d: %d x %d keys, 8 heads: %.3f ms per call (pack + kernel), %.1f TFLOP/s algorithmic, %.2f G exp2/s (MUFU bound 4500)"
          % (planes, n, n, ms, fl / ms / 1e9, n * n * 8 / ms / 1e6))
print("flag", ops.igemm_error_flag())
