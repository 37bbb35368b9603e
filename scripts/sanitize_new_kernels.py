"""Small launches of the kernels added in round 2, for compute-sanitizer (memcheck / racecheck)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dreg_nerf_b200 as pkg  # noqa: F401
from importlib import import_module
ops = import_module("dreg-nerf_b200.ops")
dev = torch.device("cuda:0")
torch.manual_seed(0)
which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("att", "all"):
    for planes, n in ((1, 700), (2, 300)):
        qkv = torch.randn(2 * n - 37, 768, device=dev)
        ops.mha_tc(qkv, n, [(-1, 0)], planes=planes)
        ops.mha_tc(qkv, n, [(0, 1), (1, 1)], planes=planes)
if which in ("elem", "all"):
    x = torch.randn(2, 512, 256, device=dev)
    ga, be, rm, rv = torch.ones(256, device=dev), torch.zeros(256, device=dev), torch.zeros(256, device=dev), torch.ones(256, device=dev)
    ops.batchnorm_small(x, ga, be, rm, rv, True, residual=x.clone(), relu=True, want_planes=True, want_stats=True)
    ops.batchnorm_fused(torch.randn(2, 4096, 64, device=dev), ga[:64], be[:64], rm[:64], rv[:64], True, want_planes=True)
    base = torch.randn(2, 9, 9, 9, 128, device=dev)
    ops.im2col(base.permute(0, 4, 3, 1, 2), 3, 2, 1, 27 * 128)
    ops.maxpool3d(torch.randn(2, 9, 8, 7, 64, device=dev))
    A, B, Cm = torch.randn(8, 333, 257, device=dev), torch.randn(8, 257, 32, device=dev), torch.zeros(8, 333, 32, device=dev)
    ops.sgemm_strided(A, (333 * 257, 257, 1), B, (257 * 32, 32, 1), Cm, (333 * 32, 32), 333, 32, 257, batch=8)
    At = torch.randn(8, 257, 333, device=dev)
    ops.sgemm_strided(At, (333 * 257, 1, 333), B, (257 * 32, 32, 1), Cm, (333 * 32, 32), 333, 32, 257, batch=8)
if which in ("wgrad", "all"):
    # staged vector reductions + transposition, taps side by side in the N tile (clamped tail), im2col unpacking
    for planes in (1, 2):
        for (g, d, h, w, cin, cout, k) in ((1, 8, 16, 32, 128, 136, 3), (2, 16, 16, 16, 64, 64, 3), (1, 1, 1, 300, 256, 72, 1)):
            x = torch.randn(g * d * h * w, cin, device=dev)
            dy = torch.randn(g * d * h * w, cout, device=dev)
            xh, xl = ops.split_planes(x, want_lo=planes == 2)
            dyp = ops.grad_split(dy, pair=planes == 2)
            xh = xh.view(g, d, h, w, cin)
            xl = xl.view(g, d, h, w, cin) if xl is not None else None
            for stage in (True, False):
                ops.conv3d_wgrad(dyp, (xh, xl), k, cout, cin, planes=planes, stage=stage)
    col = torch.randn(512, 512, device=dev)
    xh, xl = ops.split_planes(col)
    dyp = ops.grad_split(torch.randn(512, 64, device=dev))
    ops.conv3d_wgrad(dyp, (xh.view(1, 1, 1, 512, 512), xl.view(1, 1, 1, 512, 512)), 1, 64, 512, c_real=4, taps_real=125)
torch.cuda.synchronize()
print("flag", ops.igemm_error_flag(), "done", which)
