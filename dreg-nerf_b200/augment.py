"""Training-time augmentations of ``NeRFRegDataset`` (conerf/datasets/register/dataset.py:252,272-273,277-331)
on the DEVICE: the reference applies them on the host to two 58.7 MB grids per pair before the H2D copy; here
they run on the resident grids (a few element-wise torch launches over the K masked voxels only), which removes
that CPU work from the loop that feeds the forward.

Same arithmetic and the same quirks as the reference: the perturbation is centred on the mean of ALL voxel
coordinates of the perturbed grid, zeros of the unmasked voxels included (dataset.py:301-302); jitter scale
0.005 (:162), perturbation std 0.1 (:166).  Random draws come from one ``torch.Generator`` so that a run is
reproducible; pass explicit draws to reproduce a given host-side sample.
"""
import math

import torch


def _flat_xyz(grid):
    """[1, 7, Z, X, Y] storage -> writable [X*Y*Z, 3] view in the mask's (X, Y, Z) C order, or None when the
    memory layout does not allow a view (then the caller gathers / scatters)."""
    v = grid[:, :3].permute(0, 3, 4, 2, 1)
    try:
        return v.view(1, -1, 3)[0]
    except RuntimeError:
        return None


def _apply_masked(grid, mask, fn):
    flat = _flat_xyz(grid)
    if flat is not None:
        flat[mask] = fn(flat[mask])
        return grid
    _, _, Z, X, Y = grid.shape
    x, y, z = mask // (Y * Z), (mask // Z) % Y, mask % Z
    pts = grid[0, :3, z, x, y].t()
    grid[0, :3, z, x, y] = fn(pts).t()
    return grid


def points_jitter(grid, mask, scale=0.005, generator=None, noise=None):
    """dataset.py:277-285: Gaussian noise (std ``scale``) on the xyz of the masked voxels, in place."""
    mask = mask.to(grid.device)
    if noise is None:
        noise = torch.randn((mask.numel(), 3), generator=generator, device=grid.device, dtype=grid.dtype) * scale
    return _apply_masked(grid, mask, lambda p: p + noise.to(p))


def sample_se3_small(std=0.1, generator=None, device="cpu"):
    """dataset.py:71-91: axis uniform on the sphere, angle ~ N(0, 1) std pi / sqrt(3), translation ~ N(0, 1)^3
    std / sqrt(3) -> [4, 4] float32."""
    g = dict(generator=generator, device=device, dtype=torch.float64)
    phi = torch.rand((), **g) * 2.0 * math.pi
    cos_t = torch.rand((), **g) * 2.0 - 1.0
    sin_t = torch.sqrt(1.0 - cos_t * cos_t)
    axis = torch.stack([sin_t * torch.cos(phi), sin_t * torch.sin(phi), cos_t])
    theta = torch.randn((), **g) * std * math.pi / math.sqrt(3.0)
    k = torch.zeros((3, 3), dtype=torch.float64, device=device)
    k[0, 1], k[0, 2], k[1, 2] = -axis[2], axis[1], -axis[0]
    k = k - k.t()
    rot = torch.eye(3, dtype=torch.float64, device=device) + torch.sin(theta) * k + (1 - torch.cos(theta)) * (k @ k)
    mat = torch.eye(4, dtype=torch.float64, device=device)
    mat[:3, :3] = rot
    mat[:3, 3] = torch.randn(3, **g) * std / math.sqrt(3.0)
    return mat.float()


def rigid_perturb(data, std=0.1, generator=None, perturb=None, perturb_source=None):
    """dataset.py:287-323: a small rigid motion of the source or target points about the grid's mean
    coordinate, with the ground-truth pose updated so that tgt = pose(src) still holds."""
    dev = data["src_xyz_rgba"].device
    if perturb is None:
        perturb = sample_se3_small(std, generator, dev)
    if perturb_source is None:
        perturb_source = bool(torch.rand((), generator=generator, device=dev) > 0.5)
    perturb = perturb.to(device=dev, dtype=torch.float32)
    which = "src" if perturb_source else "tgt"
    grid = data[which + "_xyz_rgba"]
    centroid = grid[0, :3].reshape(3, -1).mean(dim=1)           # over ALL voxels, as the reference
    center = torch.eye(4, device=dev)
    center[:3, 3] = -centroid
    perturb = torch.linalg.inv(center) @ perturb @ center
    pose = data["pose"].to(dev)
    data["pose"] = pose @ torch.linalg.inv(perturb) if perturb_source else perturb @ pose
    rot, trans = perturb[:3, :3], perturb[:3, 3]
    _apply_masked(grid, data[which + "_mask"].to(dev), lambda p: p @ rot.t() + trans)
    return data


def random_swap(data, generator=None, swap=None):
    """dataset.py:325-331."""
    if swap is None:
        swap = bool(torch.rand((), generator=generator, device=data["src_xyz_rgba"].device) > 0.5)
    if swap:
        for a, b in (("src_xyz_rgba", "tgt_xyz_rgba"), ("src_nerf_path", "tgt_nerf_path"), ("src_mask", "tgt_mask")):
            if a in data and b in data:
                data[a], data[b] = data[b], data[a]
        data["pose"] = torch.linalg.inv(data["pose"])
    return data


def augment_pair(data, generator=None, scale=0.005, std=0.1):
    """The train-mode branch of NeRFRegDataset.__getitem__ (dataset.py:250-273) on an already loaded pair."""
    data["src_xyz_rgba"] = points_jitter(data["src_xyz_rgba"], data["src_mask"], scale, generator)
    data["tgt_xyz_rgba"] = points_jitter(data["tgt_xyz_rgba"], data["tgt_mask"], scale, generator)
    data = rigid_perturb(data, std, generator)
    return random_swap(data, generator)
