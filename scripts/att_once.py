"""One tcgen05 attention call (debug helper for compute-sanitizer / ncu)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dreg_nerf_b200 as pkg
from importlib import import_module
ops = import_module("dreg-nerf_b200.ops")
planes = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else 300
torch.manual_seed(0)
qkv = torch.randn(n, 768).cuda()
out = ops.mha_tc(qkv, n, [(0, 0)], planes=planes)
print("flag", ops.igemm_error_flag(), ops.error_flag_detail())
q, k, v = qkv[:, :256].double().cpu(), qkv[:, 256:512].double().cpu(), qkv[:, 512:].double().cpu()
qh = q.view(n, 8, 32).transpose(0, 1) / 32 ** 0.5
att = torch.softmax(qh @ k.view(n, 8, 32).transpose(0, 1).transpose(1, 2), -1)
ref = (att @ v.view(n, 8, 32).transpose(0, 1)).transpose(0, 1).reshape(n, 256)
print("rel err", float((out.cpu().double() - ref).abs().max() / ref.abs().max()))
