"""Mirror of ``conerf/loss/confidence_loss.py:56-160 compute_visibility_score``: the per-point overlap /
visibility targets of the registration training losses.

The reference re-loads the NeRF checkpoint from disk on every call and marches ``Nc x Np`` rays through
nerfacc in chunks of 60 000 with a tiny-cuda-nn density query per sample; here the score is one call of
the surface-field marcher (``drb_surface_mask``: hash grid + tensor-core MLP fused into the ray loop) on
an already resident field.  The scores are training *targets* (no gradient flows through them in the
reference either).
"""
import ctypes as C
from typing import List

import torch

from . import _lib


@torch.no_grad()
def surface_field_mask(radiance_field, occupancy_binary, points, cam_origins, roi_aabb, scene_aabb,
                       render_step_size, cut_off=0.5):
    """bool [N]: does any camera see a surface-field value max_t alpha(t) T(t) >= cut_off on its ray to the
    point (sample_grid.py:245-318 / confidence_loss.py:93-157)."""
    lib = _lib.load()
    pts = points.reshape(-1, 3).contiguous().float()
    if not pts.is_cuda:
        raise _lib.DrbError("libdregb200 has no CPU path: move the points to a CUDA device")
    dev = pts.device
    occ = occupancy_binary.to(device=dev, dtype=torch.uint8).contiguous()
    res = int(occ.shape[0])
    if occ.dim() != 3 or tuple(occ.shape) != (res, res, res):
        raise ValueError("occupancy grid must be cubic [R, R, R], got %s" % (tuple(occ.shape),))
    cams = cam_origins.to(device=dev, dtype=torch.float32).reshape(-1, 3).contiguous()
    roi = (C.c_float * 6)(*[float(v) for v in torch.as_tensor(roi_aabb).reshape(-1).tolist()])
    scene = (C.c_float * 6)(*[float(v) for v in torch.as_tensor(scene_aabb).reshape(-1).tolist()])
    out = torch.empty(pts.shape[0], dtype=torch.uint8, device=dev)
    ps = radiance_field._params_struct()
    with torch.cuda.device(dev):
        _lib.check(lib.drb_surface_mask(C.byref(ps), _lib.ptr(occ), res, roi, scene, _lib.ptr(pts), pts.shape[0],
                                        _lib.ptr(cams), cams.shape[0], float(render_step_size), float(cut_off),
                                        _lib.ptr(out), _lib.stream_ptr()), "drb_surface_mask")
        _lib.check_stream_flag("drb_surface_mask")
    return out.bool()


@torch.no_grad()
def compute_visibility_score(xyz_list: List[torch.Tensor], radiance_field, occupancy_binary, meta_data,
                             delta: float = 1e-2, cut_off: float = 0.5,
                             score_type: str = "surface_field") -> List[torch.Tensor]:
    """confidence_loss.py:56-160 with the NeRF passed in instead of ``nerf_model_path`` (load it once with
    ``blockio.load_field``).  ``xyz_list``: tensors [num_layers, N, 3] in the NeRF's world frame;
    ``meta_data``: ``aabb``, ``render_step_size``, ``camera_poses`` [Nc, 4, 4] as eval_ngp_nerf.py stores them.
    -> list of float tensors [num_layers, N, 1]: clip(1 - exp(-delta * density), 0, 1) for
    ``density_field``, the binary surface-field score for ``surface_field``."""
    if score_type not in ("density_field", "surface_field"):
        raise ValueError("score_type must be 'density_field' or 'surface_field'")
    from .ngp import check_march_options
    check_march_options(meta_data, cut_off)
    scores = []
    for xyz in xyz_list:
        num_layers, num_points = xyz.shape[0], xyz.shape[1]
        if score_type == "density_field":
            density = radiance_field.query_density(xyz.reshape(-1, 3))
            alpha = torch.clip(1 - torch.exp(-delta * density), 0, 1)
            scores.append(alpha.reshape(num_layers, num_points, 1))
            continue
        cams = meta_data["camera_poses"][..., :3, 3]
        aabb = meta_data["aabb"]
        m = surface_field_mask(radiance_field, occupancy_binary, xyz.reshape(-1, 3), cams, aabb, aabb,
                               meta_data["render_step_size"], cut_off)
        scores.append(m.float().reshape(num_layers, num_points, 1))
    return scores
