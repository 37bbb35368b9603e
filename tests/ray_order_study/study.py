"""Replays camera orderings / budget schedules on the per-ray table of one synthetic bench block (see README.md)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import dreg_nerf_b200 as pkg  # noqa: E402
from oracle import extract, extract_c  # noqa: E402
from oracle.make_goldens import make_field  # noqa: E402

RES, CAMS, NPICK = 128, 50, 3000
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 500
lib = C.CDLL(os.environ.get("STATS_LIB", "/tmp/libstats.so"))
occ, poses = pkg.synthetic.extract_scene(RES, CAMS)
meta = pkg.synthetic.extract_meta(poses)
_, ref = make_field(pkg, seed, 8.0)
idx = torch.nonzero(occ.flatten())[:, 0]
gen = torch.Generator().manual_seed(0)
roi = list(pkg.synthetic.AABB)
pts = extract.sample_points(idx, torch.rand(idx.numel(), 3, generator=gen), RES, roi)
pick = torch.randperm(idx.numel(), generator=gen)[:NPICK]
dens = extract_c.query_density(pts[pick], ref["aabb"], ref["table"], ref["w1"], ref["w2"])
P = pts[pick][dens > 0.7].contiguous()
n = P.shape[0]


def f32(t):
    a = np.ascontiguousarray(torch.as_tensor(t, dtype=torch.float32).numpy())
    return a, a.ctypes.data_as(C.POINTER(C.c_float))


keep = [f32(ref[k]) for k in ("table", "w1", "w2", "aabb")] + [f32(torch.tensor(roi)), f32(P), f32(poses[:, :3, 3].contiguous())]
occ_a = np.ascontiguousarray(occ.reshape(-1).to(torch.uint8).numpy())
best = np.zeros((n, CAMS), np.float32)
ns = np.zeros((n, CAMS), np.int32)
nocc = np.zeros((n, CAMS), np.int32)
fo = np.zeros((n, CAMS), np.float32)
lib.exp_stats(keep[0][1], keep[1][1], keep[2][1], keep[3][1], occ_a.ctypes.data_as(C.POINTER(C.c_uint8)), RES, keep[4][1],
              keep[4][1], keep[5][1], n, keep[6][1], CAMS, C.c_float(meta["render_step_size"]), C.c_float(0.5),
              best.ctypes.data_as(C.POINTER(C.c_float)), ns.ctypes.data_as(C.POINTER(C.c_int)),
              nocc.ctypes.data_as(C.POINTER(C.c_int)), fo.ctypes.data_as(C.POINTER(C.c_float)))
seen = best >= 0.5
any_seen = seen.any(1)


def cost(order):
    tot = rays = 0
    for i in range(n):
        for c in order[i]:
            tot += ns[i, c]
            rays += 1
            if seen[i, c]:
                break
    return tot / n, rays / n


def budget(B):
    tot = left = 0
    for i in range(n):
        hit = False
        for c in range(CAMS):
            tot += min(ns[i, c], B)
            if seen[i, c] and ns[i, c] <= B:
                hit = True
                break
        if not hit:
            left += 1
            for c in range(CAMS):
                tot += ns[i, c]
                if seen[i, c]:
                    break
    return tot / n, left / n


lower = sum((ns[i][seen[i]].min() if any_seen[i] else ns[i].sum()) for i in range(n)) / n
print("seed %d: %d dense points of %d sampled cells, seen by some camera %.3f" % (seed, n, NPICK, any_seen.mean()))
print("  samples / ray %.1f, occupied samples / ray %.1f, rays that see their point %.3f" % (ns.mean(), nocc.mean(), seen.mean()))
print("  camera index order (shipped)      : %.1f samples / point, %.2f rays / point" % cost([range(CAMS)] * n))
print("  fewest occupied cells first (DDA) : %.1f samples / point, %.2f rays / point" % cost(np.argsort(nocc, 1, kind="stable")))
print("  most occupied cells first         : %.1f samples / point, %.2f rays / point" % cost(np.argsort(-nocc, 1, kind="stable")))
for B in (2, 4, 8):
    print("  pre-pass with a budget of %d samples per ray, then the shipped march for the rest: %.1f samples / point (%.3f of the points left)" % ((B,) + budget(B)))
print("  clairvoyant lower bound (cheapest seeing ray; all rays for unseen points): %.1f samples / point, of which unseen points %.1f"
      % (lower, ns[~any_seen].sum() / n))
