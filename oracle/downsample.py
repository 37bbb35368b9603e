"""R4: hierarchical voxel-average down-sampling (CPU restatement, TEST INFRASTRUCTURE).

Follows conerf/register/grid_downsample.py:6-94.  The arithmetic of the
reference lives in MinkowskiEngine (un-vendored, unpinned git master,
scripts/env/install.sh:10-13): ``ME.utils.batched_coordinates`` floors the float
coordinates ``points / dl`` to int32 and prefixes the batch index;
``ME.SparseTensor(..., UNWEIGHTED_AVERAGE)`` keeps one row per unique
(batch, cell) and averages all 259 columns ([xyz | feat]) of the rows that fall
in it.  ME's output order is implementation defined (grid_downsample.py:8-10);
this restatement DEFINES it: rows grouped by cloud (src first, as
nerf_regtr.py:165-168 assumes), cells in ascending lexicographic (cx, cy, cz)
order, members summed in fp32 in ascending input-row order and divided by the
member count.  PARITY UNPINNED against ME itself.
"""
import torch


def subsample_dl_schedule(num_hierarchical, init_subsample_dl=0.025, subsample_radius=2.75):
    """dl for every round, computed in double exactly as grid_downsample.py:68-88."""
    radius_normal = init_subsample_dl * subsample_radius
    out = []
    for _ in range(num_hierarchical):
        out.append(2 * radius_normal / subsample_radius)
        radius_normal *= 2
    return out


def batched_grid_subsample(points, features, batched_lengths, sample_dl=0.1):
    """grid_downsample.py:6-44 with the order/arith defined in the module docstring."""
    lengths = [int(v) for v in batched_lengths]
    rows = torch.cat([points, features], dim=-1)      # fp32 in the reference; dtype-generic for fp64 studies
    dl32 = torch.tensor(sample_dl, dtype=torch.float32)
    out_rows, out_len = [], []
    start = 0
    for n in lengths:
        r = rows[start:start + n]
        start += n
        if n == 0:
            out_rows.append(r)
            out_len.append(0)
            continue
        cell = torch.floor(r[:, :3].float() / dl32).to(torch.int64)  # int32 range in ME
        uniq, inverse = torch.unique(cell, dim=0, sorted=True, return_inverse=True)
        acc = torch.zeros((uniq.shape[0], r.shape[1]), dtype=rows.dtype)
        acc.index_add_(0, inverse, r)                                # ascending row order
        cnt = torch.zeros(uniq.shape[0], dtype=rows.dtype)
        cnt.index_add_(0, inverse, torch.ones(n, dtype=rows.dtype))
        out_rows.append(acc / cnt[:, None])
        out_len.append(uniq.shape[0])
    return torch.cat(out_rows, dim=0), torch.tensor(out_len, dtype=torch.int64)


def hierarchical_grid_subsample(points, features, point_lengths, num_hierarchical=4,
                                init_subsample_dl=0.025, subsample_radius=2.75):
    """grid_downsample.py:47-94 (same early exit at <= 2 * 1500 rows, :70,:91)."""
    max_num_points = 1500
    ds_points, ds_features, ds_len = points, features, point_lengths
    for dl in subsample_dl_schedule(num_hierarchical, init_subsample_dl, subsample_radius):
        rows, ds_len = batched_grid_subsample(ds_points, ds_features, ds_len, dl)
        ds_points, ds_features = rows[..., :3], rows[..., 3:]
        if ds_points.shape[0] <= 2 * max_num_points:
            break
    return ds_points, ds_features, ds_len
