"""A3 - A6: voxel sampling, surface-field mask, voxel_grid scatter (TEST INFRASTRUCTURE).

Follows conerf/register/sample_grid.py:208-343, conerf/utils/nerfacc_utils.py:84-222 and
eval_ngp_nerf.py:337-412.  The ray marcher of the reference is nerfacc 0.3.5's CUDA kernel
(un-vendored): restated from that release's published source semantics - fixed step, AABB
contraction, empty cells skipped to the next voxel boundary in whole steps.  PARITY UNPINNED except
for the transmittance docstring vector (nerfacc_utils.py:56-63), which tests/ check.
Pure-Python ray loop: small cases only.
"""
import math

import torch

from oracle import ngp


def transmittance_from_alpha(alphas, ray_indices):
    """Exclusive per-ray product of (1 - alpha) (nerfacc render_transmittance_from_alpha)."""
    out = torch.ones_like(alphas)
    t, prev = 1.0, None
    for i in range(alphas.shape[0]):
        r = int(ray_indices[i])
        if r != prev:
            t, prev = 1.0, r
        out[i] = t
        t = t * (1.0 - float(alphas[i]))
    return out


def sample_points(indices, jitter, res, roi_aabb):
    """sample_grid.py:223-243: flat (X,Y,Z) C-order indices + U[0,1) jitter -> world points."""
    roi = torch.as_tensor(roi_aabb, dtype=torch.float32)
    coords = torch.stack([indices // (res * res), (indices // res) % res, indices % res], dim=1)
    x01 = (coords.float() + jitter) / float(res)
    return x01 * (roi[3:] - roi[:3]) + roi[:3]


def _occupied(x, occ, res, roi):
    u = (x - roi[:3]) / (roi[3:] - roi[:3])
    if not bool(((u >= 0) & (u < 1)).all()):
        return False
    idx = torch.clamp((u * res).to(torch.int64), 0, res - 1)
    return bool(occ[idx[0], idx[1], idx[2]])


def surface_mask(points, cam_origins, occ, res, roi_aabb, scene_aabb, step, cut_off, density_fn,
                 return_best=False):
    """sample_grid.py:245-318: surface[p] = any camera sees max_s(alpha_s * T_s) >= cut_off.
    ``return_best``: also return, per point, the largest surface-field value any marched ray reached
    (tests use it to set aside decisions that hang on the last bits of the density)."""
    roi = torch.as_tensor(roi_aabb, dtype=torch.float32)
    scene = torch.as_tensor(scene_aabb, dtype=torch.float32)
    out = torch.zeros(points.shape[0], dtype=torch.bool)
    best_of = torch.zeros(points.shape[0], dtype=torch.float64)
    f32 = torch.float32
    step = torch.tensor(step, dtype=f32)
    for pi in range(points.shape[0]):
        for ci in range(cam_origins.shape[0]):
            if out[pi]:
                break
            o = cam_origins[ci]
            d = points[pi] - o
            length = torch.sqrt((d * d).sum())
            if not float(length) > 0:
                continue
            d = d / length
            inv = 1.0 / d
            t0s, t1s = (scene[:3] - o) * inv, (scene[3:] - o) * inv
            tn = torch.minimum(t0s, t1s).max()
            tf = torch.maximum(t0s, t1s).min()
            if float(tn) > float(tf):
                continue
            near = torch.clamp(tn, min=0.0)
            t0 = near
            t1 = t0 + step
            tm = 0.5 * (t0 + t1)
            T, best = 1.0, 0.0
            while float(tm) < float(length):
                x = o + tm * d
                if _occupied(x, occ, res, roi):
                    sigma = float(density_fn(x[None])[0])
                    alpha = 1.0 - math.exp(-sigma * float(t1 - t0))
                    if T < 1e-4:
                        break
                    best = max(best, alpha * T)
                    if best >= cut_off:
                        break
                    T *= 1.0 - alpha
                    t0 = t1
                    t1 = t0 + step
                    tm = 0.5 * (t0 + t1)
                else:
                    u = (x - roi[:3]) / (roi[3:] - roi[:3]) * res
                    sgn = torch.sign(d)
                    td = (torch.floor(u + 0.5 + 0.5 * sgn) - u) * inv / res * (roi[3:] - roi[:3])
                    tt = tm + torch.clamp(td.min(), min=0.0)
                    while True:
                        tm = tm + step
                        if not float(tm) < float(tt):
                            break
                    t0, t1 = tm - 0.5 * step, tm + 0.5 * step
            best_of[pi] = max(float(best_of[pi]), best)
            if best >= cut_off:
                out[pi] = True
    return (out, best_of) if return_best else out


def surface_mask_vectorized(points, cam_origins, occ, res, roi_aabb, scene_aabb, step, cut_off, density_fn,
                            ray_subset=None):
    """Same algorithm as ``surface_mask`` with all rays marched in lock-step (torch CPU ops): the
    fastest CPU statement of the path we can offer as a timed baseline.  ``ray_subset``: optional
    LongTensor of flat ray ids (cam * n_points + point) to march only a bounded sample."""
    roi = torch.as_tensor(roi_aabb, dtype=torch.float32)
    scene = torch.as_tensor(scene_aabb, dtype=torch.float32)
    n, nc = points.shape[0], cam_origins.shape[0]
    rid = torch.arange(n * nc) if ray_subset is None else ray_subset
    pi, ci = rid % n, rid // n
    o = cam_origins[ci]
    d = points[pi] - o
    length = torch.sqrt((d * d).sum(-1))
    ok = length > 0
    d = d / length[:, None]
    inv = 1.0 / d
    t0s, t1s = (scene[:3] - o) * inv, (scene[3:] - o) * inv
    tn = torch.minimum(t0s, t1s).max(-1).values
    tf = torch.maximum(t0s, t1s).min(-1).values
    ok &= ~(tn > tf)
    stepf = torch.tensor(step, dtype=torch.float32)
    t0 = torch.clamp(tn, min=0.0)
    t1 = t0 + stepf
    tm = 0.5 * (t0 + t1)
    T = torch.ones_like(tm)
    best = torch.zeros_like(tm)
    active = ok & (tm < length)
    ext = roi[3:] - roi[:3]
    occ_flat = occ.reshape(-1)
    while bool(active.any()):
        a = torch.nonzero(active)[:, 0]
        x = o[a] + tm[a, None] * d[a]
        u = (x - roi[:3]) / ext
        inside = ((u >= 0) & (u < 1)).all(-1)
        idx = torch.clamp((u * res).to(torch.int64), 0, res - 1)
        is_occ = inside & occ_flat[(idx[:, 0] * res + idx[:, 1]) * res + idx[:, 2]]
        # ---- occupied: one sample ----
        so = a[is_occ]
        if so.numel():
            sigma = density_fn(x[is_occ])
            alpha = 1.0 - torch.exp(-sigma * (t1[so] - t0[so]))
            dead = T[so] < 1e-4
            contrib = torch.where(dead, torch.zeros_like(alpha), alpha * T[so])
            best[so] = torch.maximum(best[so], contrib)
            hit = best[so] >= cut_off
            T[so] = torch.where(dead | hit, T[so], T[so] * (1.0 - alpha))
            nt0 = t1[so]
            t0[so] = torch.where(dead | hit, t0[so], nt0)
            t1[so] = torch.where(dead | hit, t1[so], nt0 + stepf)
            tm[so] = torch.where(dead | hit, tm[so], 0.5 * (t0[so] + t1[so]))
            active[so[dead | hit]] = False
        # ---- empty: skip to the next voxel boundary in whole steps ----
        se = a[~is_occ]
        if se.numel():
            uu = u[~is_occ] * res
            sgn = torch.sign(d[se])
            td = (torch.floor(uu + 0.5 + 0.5 * sgn) - uu) * inv[se] / res * ext
            tt = tm[se] + torch.clamp(td.min(-1).values, min=0.0)
            nsteps = torch.clamp(torch.ceil((tt - tm[se]) / stepf), min=1.0)
            cand = tm[se] + nsteps * stepf
            # "do tm += step while tm < tt" in fp32: fix the rare off-by-one of the closed form
            cand = torch.where(cand - stepf >= tt, cand - stepf, cand)
            cand = torch.where(cand < tt, cand + stepf, cand)
            cand = torch.maximum(cand, tm[se] + stepf)
            tm[se] = cand
            t0[se], t1[se] = cand - 0.5 * stepf, cand + 0.5 * stepf
        active &= tm < length
    surf_ray = best >= cut_off
    out = torch.zeros(n, dtype=torch.bool)
    out.index_put_((pi[surf_ray],), torch.ones(int(surf_ray.sum()), dtype=torch.bool))
    return out


def extract_block(field, indices, jitter, occ, res, roi_aabb, scene_aabb, cam_origins, step,
                  density_thre=0.7, cut_off=0.5, with_surface=True, use_c_marcher=False):
    """eval_ngp_nerf.py:337-412 -> dict(points, rgb, alpha, density, density_mask, surface_mask, grid, mask)."""
    pts = sample_points(indices, jitter, res, roi_aabb)
    density, feat = ngp.query_density(pts, field["aabb"], field["table"], field["w1"], field["w2"])
    rgb = ngp.query_rgb_mean(ngp.fixed_viewing_directions(), feat, field["c1"], field["c2"], field["c3"])
    alpha = torch.clip(1 - torch.exp(-1e-2 * density), 0, 1)
    dmask = density > density_thre
    best = None
    if with_surface and use_c_marcher:
        # the C restatement: single precision + fused multiply-adds through dreg-nerf_b200/csrc/march_math.h, the
        # arithmetic the CUDA kernel compiles too (this scalar Python marcher keeps doubles)
        from oracle import extract_c
        smask, best, _ = extract_c.surface_mask(pts, cam_origins, occ, res, roi_aabb, scene_aabb, step, cut_off, field,
                                                all_rays=True)
    elif with_surface:
        dens_fn = lambda x: ngp.query_density(x, field["aabb"], field["table"], field["w1"], field["w2"])[0]
        smask = surface_mask(pts, cam_origins, occ, res, roi_aabb, scene_aabb, step, cut_off, dens_fn)
    else:
        smask = torch.ones_like(dmask)
    keep = dmask & smask
    grid = torch.zeros(res ** 3, 7)
    grid[indices[keep], :3] = pts[keep]
    grid[indices[keep], 3:6] = rgb[keep]
    grid[indices[keep], 6] = alpha[keep]
    return {"points": pts, "rgb": rgb, "alpha": alpha, "density": density, "density_mask": dmask,
            "surface_mask": smask, "surface_best": best, "grid": grid.reshape(res, res, res, 7), "mask": indices[keep]}
