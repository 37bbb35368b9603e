"""The kernels whose `ncu --set full` captures are summarised under profiles/ (round 2), at their full-size
shapes, each launched once (ncu replays a kernel ~40 times: keep the list short).
  1. wgrad_kernel  - level-1 FPN shape: g = 2, 64^3, 256 -> 256, 3^3, bf16, a 25 % tile list
  2. wgrad_kernel  - a transformer Linear: 900 tokens, 256 -> 1024
  3. att_fwd_kernel - 8192 queries x 8192 keys, 8 heads, bf16 and fp16-pair
  4. igemm_kernel as data gradient is the forward kernel (profiles/r01_v4_ncu_igemm_keys.txt)
  5. bn_bwd_reduce / bn_bwd_apply at the stem's size (2 x 64^3 x 64)
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dreg_nerf_b200 as pkg  # noqa: E402,F401
from importlib import import_module  # noqa: E402

ops = import_module("dreg-nerf_b200.ops")
dev = torch.device("cuda:0")
torch.manual_seed(0)

# 1. level-1 wgrad, bf16, sparse tile list
g, d, h, w, c = 2, 64, 64, 64, 256
x = torch.randn(g * d * h * w, c, device=dev)
dy = torch.randn(g * d * h * w, c, device=dev) * 1e-3
xh, _ = ops.split_planes(x, want_lo=False)
dyp = ops.grad_split(dy, pair=False)
box, tiles = ops.conv3d_tile_shape(g, d, h, w)
nt = tiles[0] * tiles[1] * tiles[2] * tiles[3]
pick = torch.randperm(nt)[: nt // 4].sort().values.int().to(dev)
cnt = torch.tensor([pick.numel()], dtype=torch.int32, device=dev)
ops.conv3d_wgrad(dyp, (xh.view(g, d, h, w, c), None), 3, c, c, planes=1, tile_list=pick, tile_count=cnt)
print("wgrad level-1: %d of %d tiles, %.1f GFLOP" % (pick.numel(), nt, 2.0 * pick.numel() * 128 * c * c * 27 / 1e9))
del x, dy, xh, dyp

# 2. linear wgrad
n = 900
x = torch.randn(n, 256, device=dev)
dy = torch.randn(n, 1024, device=dev) * 1e-3
xh, _ = ops.split_planes(x, want_lo=False)
dyp = ops.grad_split(dy, pair=False)
ops.conv3d_wgrad(dyp, (xh.view(1, 1, 1, n, 256), None), 1, 1024, 256, planes=1)

# 2b. layer1.conv2 wgrad: 64 -> 64, 3^3 at 2 x 32^3 (four taps side by side in one 256-wide N tile)
g, d, h, w, c = 2, 32, 32, 32, 64
x = torch.randn(g * d * h * w, c, device=dev)
dy = torch.randn(g * d * h * w, c, device=dev) * 1e-3
xh, _ = ops.split_planes(x, want_lo=False)
dyp = ops.grad_split(dy, pair=False)
ops.conv3d_wgrad(dyp, (xh.view(g, d, h, w, c), None), 3, c, c, planes=1)
del x, dy, xh, dyp

# 3. attention, 8k x 8k
n = 8192
qkv = torch.randn(2 * n, 768, device=dev)
ops.mha_tc(qkv, n, [(0, 1)], planes=1)
ops.mha_tc(qkv, n, [(0, 1)], planes=2)
print("attention: %.1f GFLOP per call (4 Nq Nk 32 x 8 heads)" % (4.0 * n * n * 32 * 8 / 1e9))

# 5. BatchNorm backward at the stem's size
gm, m, c = 2, 64 ** 3, 64
raw = torch.randn(gm, m, c, device=dev)
dyb = torch.randn(gm, m, c, device=dev)
ga, be = torch.ones(c, device=dev), torch.zeros(c, device=dev)
rm, rv = torch.zeros(c, device=dev), torch.ones(c, device=dev)
stats = ops.bn_forward_stats(raw, ga, be, rm, rv, True)
ops.bn_backward(dyb, raw, stats, ga, True, relu=True)
torch.cuda.synchronize()
print("done")
