"""Implicit-GEMM launches behind profiles/r02_ncu_igemm_keys.txt (`ncu --set full -k regex:igemm`), each once:
  1. igemm_kernel,  fp16 hi/lo pairs (fp32-grade), 256 -> 256, 3^3 at 2 x 32^3   (upsample_transform_2's shape)
  2. igemm2_kernel, bf16, the same shape (two M tiles per item)
  3. igemm2_kernel, bf16, 64 -> 256, 3^3 at 2 x 64^3                               (pyramid_transformation_1, dense)
  4. igemm_kernel,  fp16 pairs, 64 -> 64, 3^3 at 2 x 32^3                          (layer1.*.conv2: narrow N tile)
  5. igemm_kernel,  fp16 pairs, 866 tokens, 256 -> 768                             (in_proj: latency bound)
"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dreg_nerf_b200 as pkg  # noqa: E402,F401
from importlib import import_module  # noqa: E402

ops = import_module("dreg-nerf_b200.ops")
dev = torch.device("cuda:0")
torch.manual_seed(0)


def conv(g, d, h, w, cin, cout, k, pair):
    x = torch.randn(g, d, h, w, cin, device=dev)
    wt = torch.randn(cout, cin, k, k, k, device=dev) / math.sqrt(cin * k ** 3)
    xp = ops.split_planes(x.view(-1, cin), want_lo=pair)
    xp = tuple(t.view(g, d, h, w, cin) if t is not None else None for t in xp)
    wp = ops.pack_conv_weight(wt, pair=pair)
    torch.cuda.synchronize()
    ops.conv3d_igemm(xp, wp, k, planes=2 if pair else 1)
    torch.cuda.synchronize()
    print("conv %dx%dx%dx%d %d->%d k=%d %s: %.1f GFLOP" % (g, d, h, w, cin, cout, k, "pair" if pair else "bf16",
                                                           2.0 * g * d * h * w * cin * cout * k ** 3 / 1e9))


conv(2, 32, 32, 32, 256, 256, 3, True)
conv(2, 32, 32, 32, 256, 256, 3, False)
conv(2, 64, 64, 64, 64, 256, 3, False)
conv(2, 32, 32, 32, 64, 64, 3, True)
conv(1, 1, 1, 866, 256, 768, 1, True)
print("flag", ops.igemm_error_flag())
