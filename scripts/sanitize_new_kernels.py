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
torch.cuda.synchronize()
print("flag", ops.igemm_error_flag(), "done", which)
