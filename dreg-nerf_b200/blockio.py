"""On-disk producer / consumer of the registration path: the files the reference's extract writes
(``Evaluator.sample_points``, eval_ngp_nerf.py:397-412) and its dataset reads
(``NeRFRegDataset.__getitem__``, conerf/datasets/register/dataset.py:221-269), byte compatible:

    <root>/<dataset>/nerf_models/<scene>/block_<k>/voxel_grid.pt   torch.save(float32 [X, Y, Z, 7])
                                                    voxel_mask.pt   torch.save(int64 [K])
                                                    model.pth       NeRF checkpoint (not written here)

Rows of voxel_grid are (x, y, z, r, g, b, alpha), zeros outside the mask; the mask holds flat indices
over (X, Y, Z) in C order.  The loader applies the reference's ``permute(3, 2, 0, 1).unsqueeze(0)``
(-> [1, 7, Z, X, Y]) and returns the dict NeRFRegTr.forward consumes.  Host buffers are pinned and
copied with non_blocking=True so that the 2 x 58.7 MB of a 128^3 pair overlap with compute.
"""
import os

import torch

GRID_FILE, MASK_FILE, MODEL_FILE = "voxel_grid.pt", "voxel_mask.pt", "model.pth"
FEATURE_DIM = 7


def block_dir(root_fp, dataset, scene, k, model_dir="nerf_models"):
    """dataset.py:118-127: <root>/<dataset>/<model_dir>/<scene>/block_<k> (root_fp already holds the dataset)."""
    return os.path.join(root_fp, model_dir, scene, "block_" + str(k))


def save_block(directory, voxel_grid, voxel_mask):
    """Writes voxel_grid.pt / voxel_mask.pt exactly as eval_ngp_nerf.py:397-412 does."""
    if voxel_grid.dim() != 4 or voxel_grid.shape[-1] != FEATURE_DIM:
        raise ValueError("voxel_grid must be [X, Y, Z, 7], got %s" % (tuple(voxel_grid.shape),))
    if voxel_grid.dtype != torch.float32 or voxel_mask.dtype != torch.int64 or voxel_mask.dim() != 1:
        raise ValueError("voxel_grid must be float32 and voxel_mask a 1-D int64 index tensor")
    os.makedirs(directory, exist_ok=True)
    torch.save(voxel_grid, os.path.join(directory, GRID_FILE))
    torch.save(voxel_mask, os.path.join(directory, MASK_FILE))
    return os.path.join(directory, GRID_FILE), os.path.join(directory, MASK_FILE)


def load_block(directory, device=None, pin=True):
    """-> (xyz_rgba [1, 7, Z, X, Y] view of the stored grid, mask int64 [K]) as dataset.py:244-248."""
    grid = torch.load(os.path.join(directory, GRID_FILE), map_location="cpu")
    mask = torch.load(os.path.join(directory, MASK_FILE), map_location="cpu")
    if grid.dim() != 4 or grid.shape[-1] != FEATURE_DIM:
        raise ValueError("%s: expected [X, Y, Z, 7], got %s" % (directory, tuple(grid.shape)))
    if device is not None and torch.device(device).type == "cuda":
        if pin:
            grid, mask = grid.pin_memory(), mask.pin_memory()
        grid = grid.to(device, non_blocking=True)
        mask = mask.to(device, non_blocking=True)
    return grid.permute(3, 2, 0, 1).unsqueeze(dim=0), mask


def load_pair(src_dir, tgt_dir, device=None, src_transform=None, tgt_transform=None, **extras):
    """The ``data`` dict of NeRFRegDataset.__getitem__ (eval mode: no jitter / perturbation)."""
    src, src_mask = load_block(src_dir, device)
    tgt, tgt_mask = load_block(tgt_dir, device)
    data = {"src_xyz_rgba": src, "tgt_xyz_rgba": tgt, "src_mask": src_mask, "tgt_mask": tgt_mask,
            "src_nerf_path": os.path.join(src_dir, MODEL_FILE), "tgt_nerf_path": os.path.join(tgt_dir, MODEL_FILE)}
    if src_transform is not None and tgt_transform is not None:
        # ground truth relative pose from source to target (dataset.py:241-242)
        pose = torch.as_tensor(tgt_transform).float() @ torch.linalg.inv(torch.as_tensor(src_transform).float())
        data["pose"] = pose.unsqueeze(0)
    data.update(extras)
    return data


def load_field(model_path, aabb=None, device=None):
    """Reads a NeRF checkpoint written by the reference's CheckPointManager (``{'model': state_dict, ...}``
    or a bare state dict) into an NGPradianceField: tiny-cuda-nn's flat ``mlp_base.params`` /
    ``color_mlp.params`` tensors and the ``aabb`` buffer (conerf/radiance_fields/ngp.py:81,92-146)."""
    from .ngp import NGPradianceField
    ckpt = torch.load(model_path, map_location="cpu")
    sd = ckpt["model"] if isinstance(ckpt, dict) and "model" in ckpt else ckpt
    if aabb is None:
        aabb = sd["aabb"]
    field = NGPradianceField(aabb=aabb)
    own = field.state_dict()
    picked = {}
    for k, v in sd.items():
        if k in own:
            if tuple(v.shape) != tuple(own[k].shape):
                raise ValueError("%s: %s has shape %s, expected %s (the kernels are specialised for the reference's "
                                 "hash-grid / MLP configuration)" % (model_path, k, tuple(v.shape), tuple(own[k].shape)))
            picked[k] = v.float()
    missing = [k for k in ("mlp_base.params", "color_mlp.params") if k not in picked]
    if missing:
        raise KeyError("%s: missing %s" % (model_path, missing))
    field.load_state_dict(picked, strict=False)
    return field.to(device) if device is not None else field
