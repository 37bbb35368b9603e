"""A1 / A2: Instant-NGP field queries, CPU restatement (TEST INFRASTRUCTURE, PARITY UNPINNED).

The reference delegates this arithmetic to tiny-cuda-nn (un-vendored, unpinned git HEAD,
scripts/env/install.sh:21; call sites conerf/radiance_fields/ngp.py:92-146,157,182,188).  This file
restates the published algorithm (Mueller et al. 2022, tiny-cuda-nn's grid encoding /
FullyFusedMLP / SphericalHarmonics encodings) in fp32 torch:
  * 16 levels, 2 features, T = 2^19, base resolution 16, per-level scale 1.4472692012786865
  * level l: scale = 2^(l*log2 b)*16 - 1, res = ceil(scale) + 1, pos = x*scale + 0.5, trilinear
    interpolation of 8 corners; dense index while res^3 (rounded up to 8) <= T, otherwise the
    spatial hash (x*1) ^ (y*2654435761) ^ (z*805459861) mod T
  * density MLP 32 -> 64 ReLU -> 16 (no biases), density = exp(out[0] - 1) * inside(aabb)
  * colour MLP [16 SH | 15 feat | 1.0 pad] -> 64 ReLU -> 64 ReLU -> 3 sigmoid
tiny-cuda-nn evaluates in fp16; neither that nor its exact parameter layout can be checked here.
"""
import math

import numpy as np
import torch

N_LEVELS, T_SIZE, BASE_RES = 16, 1 << 19, 16
PER_LEVEL_SCALE = 1.4472692012786865
PRIMES = (1, 2654435761, 805459861)


def level_table():
    """-> list of (scale fp32, res, size, offset) exactly as the CUDA host code computes them."""
    out, off = [], 0
    for l in range(N_LEVELS):
        # evaluated in double, rounded once to fp32 (library-independent; tiny-cuda-nn uses fp32 exp2f/log2f)
        scale = np.float32(2.0 ** (l * math.log2(PER_LEVEL_SCALE)) * BASE_RES - 1.0)
        res = int(np.ceil(scale)) + 1
        n = (res ** 3 + 7) // 8 * 8
        n = min(n, T_SIZE)
        out.append((float(scale), res, n, off))
        off += n
    return out, off


def table_entries():
    return level_table()[1]


def _grid_index(g, res, size):
    """g int64 [N,3] -> int64 [N] (uint32 wrap-around arithmetic)."""
    m = 0xFFFFFFFF
    stride, index = 1, torch.zeros(g.shape[0], dtype=torch.int64)
    for d in range(3):
        if stride <= size:
            index = (index + (g[:, d] & m) * stride) & m
            stride *= res
    if size < stride:
        index = ((g[:, 0] & m) * PRIMES[0]) & m
        index = index ^ (((g[:, 1] & m) * PRIMES[1]) & m)
        index = index ^ (((g[:, 2] & m) * PRIMES[2]) & m)
    return index % size


def hash_encode(xn, table):
    """xn fp32 [N,3] in the unit cube, table fp32 [entries,2] -> [N,32]."""
    levels, _ = level_table()
    feats = []
    for scale, res, size, off in levels:
        # single rounding like tiny-cuda-nn's fmaf(scale, x, 0.5f): exact product in double
        pos = (xn.double() * float(np.float32(scale)) + 0.5).float()
        fl = torch.floor(pos)
        frac = pos - fl
        g0 = fl.to(torch.int64)
        acc = torch.zeros(xn.shape[0], 2, dtype=torch.float32)
        for c in range(8):
            g = g0.clone()
            w = torch.ones(xn.shape[0], dtype=torch.float32)
            for d in range(3):
                if c & (1 << d):
                    g[:, d] += 1
                    w = w * frac[:, d]
                else:
                    w = w * (1.0 - frac[:, d])
            idx = _grid_index(g, res, size) + off
            acc = acc + w[:, None] * table[idx]
        feats.append(acc)
    return torch.cat(feats, dim=1)


def query_density(x, aabb, table, w1, w2):
    """ngp.py:148-176 -> (density [N], feat [N,15])."""
    aabb = torch.as_tensor(aabb, dtype=torch.float32)
    xn = (x - aabb[:3]) / (aabb[3:] - aabb[:3])
    inside = ((xn > 0.0) & (xn < 1.0)).all(dim=-1)
    f = hash_encode(xn, table)
    out = torch.relu(f @ w1.t()) @ w2.t()
    density = torch.exp(out[:, 0] - 1.0) * inside
    return density, out[:, 1:16]


def sh4(d):
    x, y, z = d[:, 0], d[:, 1], d[:, 2]
    xy, xz, yz, x2, y2, z2 = x * y, x * z, y * z, x * x, y * y, z * z
    return torch.stack([
        torch.full_like(x, 0.28209479177387814), -0.48860251190291987 * y, 0.48860251190291987 * z,
        -0.48860251190291987 * x, 1.0925484305920792 * xy, -1.0925484305920792 * yz,
        0.94617469575755997 * z2 - 0.31539156525251999, -1.0925484305920792 * xz,
        0.54627421529603959 * x2 - 0.54627421529603959 * y2, 0.59004358992664352 * y * (-3.0 * x2 + y2),
        2.8906114426405538 * xy * z, 0.45704579946446572 * y * (1.0 - 5.0 * z2),
        0.3731763325901154 * z * (5.0 * z2 - 3.0), 0.45704579946446572 * x * (1.0 - 5.0 * z2),
        1.4453057213202769 * z * (x2 - y2), 0.59004358992664352 * x * (-x2 + 3.0 * y2)], dim=1)


def query_rgb(dirs, feat, c1, c2, c3):
    """ngp.py:178-193: dirs [N,3] in [-1,1], feat [N,15] -> rgb [N,3]."""
    d01 = (dirs + 1.0) / 2.0
    enc = sh4(d01 * 2.0 - 1.0)
    h = torch.cat([enc, feat, torch.ones(feat.shape[0], 1)], dim=1)       # width pad fed with 1
    h = torch.relu(h @ c1.t())
    h = torch.relu(h @ c2.t())
    return torch.sigmoid(h @ c3[:3].t())


def query_rgb_mean(viewdirs, feat, c1, c2, c3):
    """sample_grid.py:332-340: mean over the fixed view directions."""
    acc = torch.zeros(feat.shape[0], 3)
    for k in range(viewdirs.shape[0]):
        acc = acc + query_rgb(viewdirs[k].repeat(feat.shape[0], 1), feat, c1, c2, c3)
    return acc / viewdirs.shape[0]


def fixed_viewing_directions():
    """sample_grid.py:132-146 (x == y, not unit length - kept as is)."""
    dirs = []
    for phi in (math.pi / 3, 0, -math.pi):
        for k in range(6):
            theta = k * math.pi / 3
            dirs.append([math.cos(phi) * math.sin(theta), math.cos(phi) * math.sin(theta), math.sin(theta)])
    return torch.tensor(dirs, dtype=torch.float32)
