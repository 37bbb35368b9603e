"""Edge cases of the path on the GPU: empty and degenerate inputs, limits, ragged clouds.  The reference has no
tests of its own here; these pin what the C ABI does instead of crashing or corrupting memory."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _ops():
    from importlib import import_module
    return import_module("dreg-nerf_b200.ops")


def _meta(cams):
    poses = torch.eye(4).repeat(cams.shape[0], 1, 1)
    poses[:, :3, 3] = cams
    return {"aabb": [-1.5] * 3 + [1.5] * 3, "render_step_size": 3.0 * math.sqrt(3) / 1024,
            "cone_angle": 0.0, "alpha_thre": 0.0, "camera_poses": poses}


def test_extract_of_an_empty_block(pkg, cuda):
    """No candidate cell at all, and candidate cells none of which is dense: zero grid, empty mask, no error."""
    res = 32
    cams = torch.tensor([[4.0, 0.0, 1.0], [0.0, 4.0, 1.0]])
    sg = pkg.SampleGrid([-1.5] * 3 + [1.5] * 3, res)
    dense_field = pkg.synthetic.make_ngp_field(seed=3, table_std=8.0).to(cuda)
    grid, mask = pkg.extract_block(dense_field, sg, torch.zeros(res, res, res, dtype=torch.bool, device=cuda), _meta(cams), cuda)
    assert mask.numel() == 0 and mask.dtype == torch.int64 and float(grid.abs().max()) == 0.0
    # tiny-cuda-nn's default initialisation: density = exp(-1) < 0.7 everywhere -> nothing passes
    thin = pkg.synthetic.make_ngp_field(seed=3, table_std=None).to(cuda)
    occ = torch.zeros(res, res, res, dtype=torch.bool)
    occ[8:24, 8:24, 8:24] = True
    grid, mask = pkg.extract_block(thin, sg, occ.to(cuda), _meta(cams), cuda)
    assert mask.numel() == 0 and float(grid.abs().max()) == 0.0
    pts, rgb, alpha, idx, dmask, smask = sg.query_radiance_and_density_from_camera(thin, occ.to(cuda), _meta(cams), cuda)
    assert pts.shape == (16 ** 3, 3) and not bool(dmask.any())
    assert torch.allclose(alpha, torch.full_like(alpha, 1 - math.exp(-0.01 * math.exp(-1.0))), atol=1e-5)


def test_surface_mask_without_points_or_cameras(pkg, cuda):
    f = pkg.synthetic.make_ngp_field(seed=4).to(cuda)
    occ = torch.ones(16, 16, 16, dtype=torch.bool, device=cuda)
    roi = [-1.5] * 3 + [1.5] * 3
    step = 3.0 * math.sqrt(3) / 1024
    none = pkg.surface_field_mask(f, occ, torch.zeros(0, 3, device=cuda), torch.tensor([[4.0, 0, 1]]), roi, roi, step)
    assert none.shape == (0,) and none.dtype == torch.bool
    blind = pkg.surface_field_mask(f, occ, torch.rand(100, 3, device=cuda) - 0.5, torch.zeros(0, 3), roi, roi, step)
    assert blind.shape == (100,) and not bool(blind.any())
    # a camera that coincides with the point: zero-length ray, nothing to march
    p = torch.tensor([[0.25, 0.25, 0.25]], device=cuda)
    same = pkg.surface_field_mask(f, occ, p, p.cpu(), roi, roi, step)
    assert same.tolist() == [False]


def test_limits_are_refused_not_overrun(pkg, cuda):
    """More cameras than the ray word encodes, a field on the CPU, a non-cubic occupancy grid."""
    f = pkg.synthetic.make_ngp_field(seed=4).to(cuda)
    occ = torch.ones(16, 16, 16, dtype=torch.bool, device=cuda)
    roi = [-1.5] * 3 + [1.5] * 3
    pts = torch.rand(10, 3, device=cuda)
    with pytest.raises(pkg.DrbError, match="cameras"):
        pkg.surface_field_mask(f, occ, pts, torch.rand(1024, 3) + 3.0, roi, roi, 0.005)
    with pytest.raises(pkg.DrbError, match="no CPU path"):
        pkg.surface_field_mask(pkg.synthetic.make_ngp_field(seed=4), occ, pts, torch.rand(2, 3) + 3.0, roi, roi, 0.005)
    with pytest.raises(ValueError, match="cubic"):
        pkg.surface_field_mask(f, occ[:, :, :8], pts, torch.rand(2, 3) + 3.0, roi, roi, 0.005)


def test_downsample_degenerate_clouds(pkg, cuda):
    """One cloud empty, one point per cloud, all points in one cell (grid_downsample.py:6-94 semantics)."""
    ops = _ops()
    ld = 260
    gen = torch.Generator().manual_seed(0)

    def rows_of(n):
        r = torch.zeros(n, ld)
        r[:, :3] = torch.rand(n, 3, generator=gen) * 2 - 1
        r[:, 4:] = torch.randn(n, 256, generator=gen)
        return r

    from oracle.downsample import hierarchical_grid_subsample
    for n_src, n_tgt in ((1, 1), (0, 40), (40, 0), (4000, 1)):
        rows = rows_of(n_src + n_tgt)
        out, a, b = ops.hierarchical_downsample(rows.to(cuda), n_src, n_tgt)
        p_ref, f_ref, l_ref = hierarchical_grid_subsample(rows[:, :3].contiguous(), rows[:, 4:].contiguous(),
                                                          torch.tensor([n_src, n_tgt]), 6)
        assert [a, b] == l_ref.tolist() and a + b == out.shape[0]
        assert (a > 0) == (n_src > 0) and (b > 0) == (n_tgt > 0)
        assert torch.equal(out[:, :3].cpu(), p_ref) and torch.equal(out[:, 4:].cpu(), f_ref)
    # every point in ONE cell of the first round: the mean of all rows
    rows = rows_of(50)
    rows[:, :3] = 0.01 + 0.001 * torch.rand(50, 3, generator=gen)
    out, a, b = ops.hierarchical_downsample(rows.to(cuda), 50, 0, num_rounds=1, max_total=1)
    assert (a, b) == (1, 0)
    assert torch.allclose(out[0].cpu(), rows.mean(0), atol=1e-5)


def test_forward_with_an_empty_mask_raises(pkg, cuda):
    """The reference would soft-max over zero keys; here the engine refuses the pair with a message."""
    torch.manual_seed(0)
    model = pkg.NeRFRegTr().to(cuda).eval()
    good = pkg.synthetic.to_device(pkg.synthetic.make_pair(res=32, pair_id=0), cuda)
    bad = dict(good, src_mask=good["src_mask"][:0])
    with torch.no_grad():
        with pytest.raises(pkg.DrbError):
            model(bad)
        out = model(dict(good))                      # and keeps working
    assert torch.isfinite(out["pose"]).all()


def test_ragged_pair_tiny_against_large_cloud(pkg, cuda):
    """A handful of source voxels against thousands of target voxels (ragged attention / Procrustes)."""
    from oracle import regtr
    torch.manual_seed(0)
    model = pkg.NeRFRegTr()
    sd = pkg.synthetic.seeded_state_dict(model, seed=0, attn_gain=4.0)
    model.load_state_dict(sd)
    model = model.to(cuda).eval()
    data = pkg.synthetic.make_pair(res=32, pair_id=2)
    data["src_mask"] = data["src_mask"][::97].contiguous()         # ~25 voxels
    with torch.no_grad():
        out = model(pkg.synthetic.to_device(dict(data), cuda))
        ref = regtr.forward(sd, dict(data), training=False)
    assert out["src_kp"][0].shape == ref["src_kp"][0].shape and out["src_kp"][0].shape[0] < 40
    rel = lambda a, b: ((a.cpu().double() - b.double()).abs().max() / b.double().abs().max()).item()
    assert rel(out["src_feats"][0], ref["src_feats"][0]) < 1e-3
    assert rel(out["tgt_feats"][0], ref["tgt_feats"][0]) < 1e-3
    assert rel(out["pose"], ref["pose"]) < 1e-3


def test_surface_mask_beyond_the_ray_word(pkg, cuda):
    """More than 2^22 points in one call (a 256^3 block with over a quarter of its cells occupied): marched in
    chunks of points, same answer as two separate calls."""
    n = (1 << 22) + 12345
    gen = torch.Generator().manual_seed(7)
    pts = (torch.rand(n, 3, generator=gen) * 2.0 - 1.0).to(cuda)
    res = 32
    ax = (torch.arange(res, dtype=torch.float32) + 0.5) / res * 3.0 - 1.5
    X, Y, Z = torch.meshgrid(ax, ax, ax, indexing="ij")
    occ = ((X ** 2 + Y ** 2 + Z ** 2) < 0.6 ** 2).to(cuda)
    f = pkg.synthetic.make_ngp_field(seed=501, table_std=8.0).to(cuda)
    roi = [-1.5] * 3 + [1.5] * 3
    cams = torch.tensor([[4.0, 0.0, 1.0], [-4.0, 0.5, 0.0]])
    step = 3.0 * math.sqrt(3) / 1024
    whole = pkg.surface_field_mask(f, occ, pts, cams, roi, roi, step)
    half = n // 2
    parts = torch.cat([pkg.surface_field_mask(f, occ, pts[:half], cams, roi, roi, step),
                       pkg.surface_field_mask(f, occ, pts[half:], cams, roi, roi, step)])
    assert whole.shape == (n,) and torch.equal(whole, parts)
    frac = float(whole.float().mean())
    print("points seen: %.3f of %d" % (frac, n))
    assert 0.0 < frac < 1.0
