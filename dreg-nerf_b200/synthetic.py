"""Seeded synthetic inputs for tests and bench.py (no dataset, no checkpoint, no network).

Register inputs follow SURVEY.md section 8d: a thin ellipsoid-shell occupancy inside the
world AABB [-1.5, 1.5]^3, one jittered point per occupied cell (as sample_grid.py:226-229),
random colour / alpha, the target being the source geometry moved by a random SE(3)
(rotation ~ N(0, sigma) axis-angle, translation clamped to +-0.2 like
conerf/geometry/pose_util.py:363-368) and re-voxelised.  Tensors come in the on-disk layout of
eval_ngp_nerf.py:397-412: grid float32 [X, Y, Z, 7] and int64 flat indices.
"""
import math

import torch

AABB = (-1.5, -1.5, -1.5, 1.5, 1.5, 1.5)


def _rodrigues(v):
    theta = float(torch.linalg.norm(v))
    if theta < 1e-12:
        return torch.eye(3, dtype=torch.float64)
    k = (v / theta).double()
    K = torch.tensor([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]], dtype=torch.float64)
    return torch.eye(3, dtype=torch.float64) + math.sin(theta) * K + (1 - math.cos(theta)) * (K @ K)


def random_se3(gen, rot_sigma=0.3):
    rot = _rodrigues(torch.randn(3, generator=gen, dtype=torch.float64) * rot_sigma)
    trans = torch.clamp(torch.randn(3, generator=gen, dtype=torch.float64) * 0.2, -0.2, 0.2)
    T = torch.eye(4, dtype=torch.float64)
    T[:3, :3], T[:3, 3] = rot, trans
    return T


def _shell_points(res, gen, n_shells=3):
    """World-space surface samples of a few random ellipsoid shells, dense enough to hit every
    shell voxel at resolution ``res``."""
    pts = []
    n = int(6 * res * res)
    for _ in range(n_shells):
        centre = (torch.rand(3, generator=gen, dtype=torch.float64) - 0.5) * 0.8
        radii = 0.35 + 0.5 * torch.rand(3, generator=gen, dtype=torch.float64)
        d = torch.randn(n, 3, generator=gen, dtype=torch.float64)
        d = d / d.norm(dim=1, keepdim=True)
        pts.append(centre + d * radii)
    return torch.cat(pts)


def voxelise(points_world, res, gen, crop_axis=None, crop_keep=1.0):
    """-> (grid [R,R,R,7] float32, mask int64 [K]) in the voxel_grid.pt / voxel_mask.pt layout."""
    lo = torch.tensor(AABB[:3], dtype=torch.float64)
    hi = torch.tensor(AABB[3:], dtype=torch.float64)
    u = (points_world - lo) / (hi - lo)
    keep = ((u > 0) & (u < 1)).all(dim=1)
    if crop_axis is not None:
        keep &= u[:, crop_axis] < crop_keep
    cell = torch.floor(u[keep] * res).long().clamp_(0, res - 1)
    flat = torch.unique(cell[:, 0] * res * res + cell[:, 1] * res + cell[:, 2])
    k = flat.numel()
    coords = torch.stack([flat // (res * res), (flat // res) % res, flat % res], dim=1).double()
    xyz01 = (coords + torch.rand(k, 3, generator=gen, dtype=torch.float64)) / res
    xyz = xyz01 * (hi - lo) + lo
    rgb = torch.rand(k, 3, generator=gen, dtype=torch.float64)
    alpha = torch.rand(k, 1, generator=gen, dtype=torch.float64)
    grid = torch.zeros(res * res * res, 7, dtype=torch.float32)
    grid[flat] = torch.cat([xyz, rgb, alpha], dim=1).float()
    return grid.reshape(res, res, res, 7), flat


def make_pair(res=32, pair_id=0, overlap_keep=0.8):
    """One (src, tgt) pair in the dict layout NeRFRegDataset.__getitem__ produces
    (conerf/datasets/register/dataset.py:244-269), CPU tensors."""
    gen = torch.Generator().manual_seed(1000 + 2 * pair_id)
    surf = _shell_points(res, gen)
    T = random_se3(gen)
    moved = surf @ T[:3, :3].T + T[:3, 3]
    src_grid, src_mask = voxelise(surf, res, gen, crop_axis=0, crop_keep=overlap_keep)
    gen_t = torch.Generator().manual_seed(1000 + 2 * pair_id + 1)
    tgt_grid, tgt_mask = voxelise(moved, res, gen_t, crop_axis=1, crop_keep=overlap_keep)
    return {
        "src_xyz_rgba": src_grid.permute(3, 2, 0, 1).unsqueeze(0),
        "tgt_xyz_rgba": tgt_grid.permute(3, 2, 0, 1).unsqueeze(0),
        "src_mask": src_mask, "tgt_mask": tgt_mask,
        "src_nerf_path": "", "tgt_nerf_path": "",
        "pose": T.float().unsqueeze(0),          # [1, 4, 4] as NeRFRegDataset builds it (dataset.py:242)
        "scene": "synthetic_%d" % pair_id, "dataset": "synthetic", "index": pair_id,
    }


def to_device(data, device):
    """conerf/utils/utils.py:29 all_to_device equivalent for the dict above."""
    return {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in data.items()}


def seeded_state_dict(module, seed=0, attn_gain=1.0):
    """Deterministic weights for any module with the NeRFRegTr key layout, independent of the
    constructor's RNG use: every floating tensor is re-drawn from a CPU generator in state_dict
    order.  Conv / linear weights ~ N(0, gain / sqrt(fan_in)); norm scales near 1; running_var
    positive.  ``attn_gain`` scales the decoder's q/k projections (sharper correspondences)."""
    gen = torch.Generator().manual_seed(seed)
    seen = {}
    out = {}
    for name, t in module.state_dict().items():
        key = (t.data_ptr(), tuple(t.shape))
        if key in seen:                       # alias entries (fpn3d.feature_pyramid.resnet.*)
            out[name] = out[seen[key]]
            continue
        seen[key] = name
        if not t.is_floating_point():
            out[name] = torch.zeros_like(t, device="cpu")
            continue
        shape = tuple(t.shape)
        leaf = name.rsplit(".", 1)[-1]
        if name.endswith("running_var"):
            v = 0.5 + torch.rand(shape, generator=gen)
        elif name.endswith("running_mean"):
            v = 0.1 * torch.randn(shape, generator=gen)
        elif t.dim() >= 2:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            v = torch.randn(shape, generator=gen) / math.sqrt(fan_in)
            if "correspondence_decoder.q_proj" in name or "correspondence_decoder.k_proj" in name:
                v = v * attn_gain
        elif leaf == "weight":                # BatchNorm / LayerNorm scale
            v = 1.0 + 0.1 * torch.randn(shape, generator=gen)
        else:                                 # biases
            v = 0.05 * torch.randn(shape, generator=gen)
        out[name] = v.float()
    return out


def make_ngp_field(seed=0, table_std=8.0):
    """Seeded random-weight Instant-NGP field on the world AABB.  tiny-cuda-nn's default init gives
    density ~ e^-1 < 0.7 everywhere (empty mask), so the hash table is drawn from N(0, table_std):
    table_std ~ 8 yields a few percent of samples dense enough to pass both the 0.7 density
    threshold and the surface-field test at the reference's step size."""
    from .ngp import NGPradianceField
    torch.manual_seed(seed)
    f = NGPradianceField(aabb=list(AABB))
    f.reset_parameters(table_std=table_std)
    return f


def extract_scene(res, n_cam):
    """Binary occupancy (an ellipsoid shell, ~8 % of the cells) and camera-to-world poses on a ring of
    radius 4 looking at the origin (only the camera centres matter for the surface-field mask)."""
    ax = (torch.arange(res, dtype=torch.float32) + 0.5) / res * 3.0 - 1.5
    X, Y, Z = torch.meshgrid(ax, ax, ax, indexing="ij")
    r = torch.sqrt(X ** 2 + (Y * 1.2) ** 2 + (Z * 0.9) ** 2)
    occ = (r < 0.9) & (r > 0.55)
    ang = torch.arange(n_cam, dtype=torch.float32) * (2 * math.pi / n_cam)
    poses = torch.eye(4).repeat(n_cam, 1, 1)
    poses[:, 0, 3] = 4 * torch.cos(ang)
    poses[:, 1, 3] = 4 * torch.sin(ang)
    poses[:, 2, 3] = 1.0
    return occ, poses


def extract_meta(poses, render_n_samples=1024):
    """meta_data dict of eval_ngp_nerf.py (aabb, render_step_size as train_ngp_nerf.py:88-92)."""
    ext = max(AABB[3] - AABB[0], AABB[4] - AABB[1], AABB[5] - AABB[2])
    return {"aabb": list(AABB), "render_step_size": ext * math.sqrt(3) / render_n_samples,
            "cone_angle": 0.0, "alpha_thre": 0.0, "camera_poses": poses}
