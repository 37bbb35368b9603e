"""Host-side logic (CPU): drop-in surface of NeRFRegTr, synthetic generator, error behaviour."""
import os

import pytest
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_state_dict_layout_matches_reference(pkg):
    """772 entries, the reference's names / shapes / order (conerf/register/nerf_regtr.py + alias)."""
    torch.manual_seed(0)
    sd = pkg.NeRFRegTr().state_dict()
    with open(os.path.join(GOLDEN, "state_dict_keys.txt")) as fh:
        want = [l.split() for l in fh.read().strip().splitlines()]
    got = [[k, "x".join(str(s) for s in v.shape)] for k, v in sd.items()]
    assert len(got) == 772
    assert got == [[w[0], w[1] if len(w) > 1 else ""] for w in want]
    assert sd["fpn3d.backbone_net.conv1.weight"].data_ptr() == sd["fpn3d.feature_pyramid.resnet.conv1.weight"].data_ptr()
    assert sum(p.numel() for p in pkg.NeRFRegTr().parameters()) == 61124225


def test_same_seed_same_init_as_reference(pkg):
    from oracle.ref_shim import import_reference, reference_available
    if not reference_available():
        pytest.skip("reference tree not present on this machine")
    ref = import_reference()
    torch.manual_seed(123)
    a = ref.NeRFRegTr().state_dict()
    torch.manual_seed(123)
    b = pkg.NeRFRegTr().state_dict()
    assert all(torch.equal(a[k], b[k]) for k in a)


def test_constructor_contract(pkg):
    m = pkg.NeRFRegTr(pos_emb_type="sine", pos_emb_dim=256, pos_emb_scaling=1.0, num_downsample=6)
    assert m.num_downsample == 6
    with pytest.raises(NotImplementedError):
        pkg.NeRFRegTr(pos_emb_type="learned")
    with pytest.raises(ValueError):
        pkg.NeRFRegTr(precision="fp8")


def test_forward_refuses_cpu_and_bad_shapes(pkg):
    m = pkg.NeRFRegTr()
    data = pkg.synthetic.make_pair(res=16, pair_id=0)
    with pytest.raises(pkg.DrbError, match="no CPU path"):
        m(dict(data))
    bad = dict(data)
    bad["src_xyz_rgba"] = data["src_xyz_rgba"][0]
    with pytest.raises(AssertionError):
        m(bad)


def test_six_dim_inputs_are_squeezed_in_place(pkg):
    """nerf_regtr.py:121-128: a leading singleton dimension is removed from the caller's dict."""
    m = pkg.NeRFRegTr()
    d = pkg.synthetic.make_pair(res=16, pair_id=0)
    d6 = {k: (v.unsqueeze(0) if torch.is_tensor(v) else [v]) for k, v in d.items() if k not in ("scene", "dataset", "index")}
    with pytest.raises(pkg.DrbError):
        m(d6)
    assert d6["src_xyz_rgba"].dim() == 5 and d6["src_mask"].dim() == 1 and d6["pose"].dim() == 3


def test_synthetic_pairs_are_deterministic_and_in_layout(pkg):
    a, b = pkg.synthetic.make_pair(res=32, pair_id=4), pkg.synthetic.make_pair(res=32, pair_id=4)
    assert torch.equal(a["src_xyz_rgba"], b["src_xyz_rgba"]) and torch.equal(a["tgt_mask"], b["tgt_mask"])
    g = a["src_xyz_rgba"]
    assert g.shape == (1, 7, 32, 32, 32) and a["src_mask"].dtype == torch.int64
    flat = g.permute(0, 3, 4, 2, 1).reshape(-1, 7)          # (X, Y, Z) C-order rows
    nz = torch.nonzero(flat.abs().sum(dim=1))[:, 0]
    assert torch.equal(nz, a["src_mask"])                    # zeros outside the mask, eval_ngp_nerf.py:397-408
    xyz = flat[a["src_mask"], :3]
    cell = torch.floor((xyz + 1.5) / 3.0 * 32).long().clamp(0, 31)
    assert torch.equal(cell[:, 0] * 1024 + cell[:, 1] * 32 + cell[:, 2], a["src_mask"])
    frac = a["src_mask"].numel() / 32 ** 3
    assert 0.005 < frac < 0.2


def test_seeded_state_dict_is_reproducible(pkg):
    m = pkg.NeRFRegTr()
    a, b = pkg.synthetic.seeded_state_dict(m, 5), pkg.synthetic.seeded_state_dict(m, 5)
    assert all(torch.equal(a[k], b[k]) for k in a)
    m.load_state_dict(a)
    assert (a["fpn3d.backbone_net.bn1.running_var"] > 0).all()


def test_reciprocal_fma_division_is_correctly_rounded():
    """The marcher replaces the IEEE divisions by the ROI extent (nerfacc roi_to_unit) with
    q = a*y, r = fma(-b, q, a), q + r*y, y = RN(1/b) (csrc/ngp.cu div_rn_rcp).  Checked here against
    exact rational arithmetic: the result must equal RN(a / b) bit for bit."""
    from fractions import Fraction
    import numpy as np

    def rn32(fr):
        f = np.float32(float(fr))
        best = None
        for c in (f, np.nextafter(f, np.float32(np.inf)), np.nextafter(f, np.float32(-np.inf))):
            d = abs(Fraction(float(c)) - fr)
            even = (int(np.float32(c).view(np.uint32)) & 1) == 0
            if best is None or d < best[0] or (d == best[0] and even):
                best = (d, c)
        return np.float32(best[1])

    def fma(a, b, c):
        return rn32(Fraction(float(a)) * Fraction(float(b)) + Fraction(float(c)))

    rng = np.random.default_rng(0)
    for b in (3.0, 2.5, 1.7, 0.3, 10.0, 3.2):
        b = np.float32(b)
        y = rn32(Fraction(1) / Fraction(float(b)))
        assert y == np.float32(1.0 / float(b))                  # what the host computes
        for a in np.concatenate([rng.uniform(-4, 4, 700), rng.uniform(-1e-3, 1e-3, 100)]).astype(np.float32):
            q = rn32(Fraction(float(a)) * Fraction(float(y)))
            r = fma(-b, q, a)
            assert fma(r, y, q) == rn32(Fraction(float(a)) / Fraction(float(b)))


def test_block_files_round_trip(tmp_path, pkg):
    """voxel_grid.pt / voxel_mask.pt written by blockio.save_block are what the reference's dataset reads
    (dataset.py:244-248): same tensors, same permute, and the mask addresses the (X, Y, Z) C-order rows."""
    import os
    import torch
    g = torch.Generator().manual_seed(3)
    grid = torch.zeros(16, 16, 16, 7)
    mask = torch.randperm(16 ** 3, generator=g)[:200].sort().values
    grid.reshape(-1, 7)[mask] = torch.rand(200, 7, generator=g)
    d0, d1 = str(tmp_path / "block_0"), str(tmp_path / "block_1")
    pkg.blockio.save_block(d0, grid, mask)
    pkg.blockio.save_block(d1, grid.flip(0).contiguous(), mask)
    assert sorted(os.listdir(d0)) == ["voxel_grid.pt", "voxel_mask.pt"]
    # the reference's reader, verbatim
    ref_xyz_rgba = torch.load(os.path.join(d0, "voxel_grid.pt")).permute(3, 2, 0, 1).unsqueeze(dim=0)
    ref_mask = torch.load(os.path.join(d0, "voxel_mask.pt"))
    data = pkg.blockio.load_pair(d0, d1, src_transform=torch.eye(4), tgt_transform=torch.eye(4), scene="s")
    assert torch.equal(data["src_xyz_rgba"], ref_xyz_rgba) and torch.equal(data["src_mask"], ref_mask)
    assert data["src_xyz_rgba"].shape == (1, 7, 16, 16, 16) and data["pose"].shape == (1, 4, 4)
    assert data["src_nerf_path"].endswith("block_0/model.pth") and data["scene"] == "s"
    # rows selected the way NeRFRegTr.forward does (nerf_regtr.py:144-147) are the stored rows
    rows = data["src_xyz_rgba"].permute(0, 3, 4, 2, 1).reshape(1, -1, 7)[0, data["src_mask"]]
    assert torch.equal(rows, grid.reshape(-1, 7)[mask])
    import pytest
    with pytest.raises(ValueError):
        pkg.blockio.save_block(d0, grid.permute(3, 0, 1, 2), mask)


def test_field_checkpoint_round_trip(tmp_path, pkg):
    """A checkpoint in the CheckPointManager layout ({'model': state_dict}) with tiny-cuda-nn's flat
    params loads into NGPradianceField; a wrong-size table is rejected."""
    import torch
    import pytest
    f = pkg.NGPradianceField(aabb=[-1.5] * 3 + [1.5] * 3)
    f.reset_parameters(table_std=2.0)
    path = str(tmp_path / "model.pth")
    torch.save({"model": f.state_dict(), "step": 7}, path)
    g = pkg.blockio.load_field(path)
    assert torch.equal(g.mlp_base.params, f.mlp_base.params) and torch.equal(g.color_mlp.params, f.color_mlp.params)
    assert torch.equal(g.aabb, f.aabb)
    bad = dict(f.state_dict())
    bad["mlp_base.params"] = bad["mlp_base.params"][:-8]
    torch.save(bad, path)
    with pytest.raises(ValueError):
        pkg.blockio.load_field(path)


def test_training_losses_match_reference_fixture(pkg):
    """InfoNCELoss against the reference's own module (fixture: oracle/make_goldens.py losses_case), value and
    gradients bit for bit on the CPU; CorrespondenceLoss against the closed form of the robust loss at the call
    site's alpha = 1, scale = 0.5."""
    import os
    import torch
    fix = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "losses.pt"))
    nce = pkg.InfoNCELoss(256, 0.2, 0.4)
    assert list(nce.state_dict().keys()) == ["W"]                     # the reference checkpoint's key
    nce.load_state_dict({"W": fix["W"]})
    sf, tf = fix["sf"].clone().requires_grad_(True), fix["tf"].clone().requires_grad_(True)
    loss = nce([sf], [tf], [fix["sx"]], [fix["tx"]])
    gs, gt, gw = torch.autograd.grad(loss, [sf, tf, nce.W])
    assert torch.allclose(loss, fix["infonce"], rtol=1e-6)
    for got, want in ((gs, fix["g_sf"]), (gt, fix["g_tf"]), (gw, fix["g_W"])):
        assert (got - want).abs().max() <= 1e-6 * want.abs().max()
    corr = pkg.CorrespondenceLoss()([fix["kp"]], [fix["pred"]], fix["pose"], overlap_weights=[fix["w"]])
    assert abs(float(corr) - float(fix["corr"])) < 1e-5 * abs(float(fix["corr"]))
    assert torch.allclose(pkg.losses.se3_inv(fix["pose"]), fix["pose_inv"])
    # the reference's broadcast quirk: [num_layers, N, 1] weights against the [N] error -> sum_j err_j
    ones = pkg.CorrespondenceLoss()([fix["kp"]], [fix["pred"]], fix["pose"], overlap_weights=[torch.ones(6, 50, 1)])
    direct = pkg.losses.robust_charbonnier(fix["pred"] - pkg.losses.se3_transform_list(fix["pose"], [fix["kp"]])[0])
    assert abs(float(ones) - float(direct.abs().sum())) < 1e-4


def test_device_side_augmentations_keep_the_pair_consistent(pkg):
    """dataset.py:277-331 restated for resident tensors: after jitter + rigid perturbation + swap the ground-truth
    pose still maps the source points of overlapping voxels onto the target frame exactly as before (the
    augmentations only move points rigidly and update `pose` accordingly); the jitter touches masked voxels only."""
    import torch
    gen = torch.Generator().manual_seed(3)
    base = pkg.synthetic.make_pair(res=16, pair_id=1)
    clone = lambda d: {k: (v.clone() if torch.is_tensor(v) else v) for k, v in d.items()}

    def masked_xyz(d, side):
        g = d[side + "_xyz_rgba"]
        return g[:, :3].permute(0, 3, 4, 2, 1).reshape(1, -1, 3)[0, d[side + "_mask"]]

    # jitter: only masked voxels move, by ~scale
    d = clone(base)
    d["src_xyz_rgba"] = pkg.augment.points_jitter(d["src_xyz_rgba"], d["src_mask"], 0.005, gen)
    diff = (d["src_xyz_rgba"] - base["src_xyz_rgba"]).abs()
    assert float(diff[:, 3:].max()) == 0.0
    moved = diff[:, :3].permute(0, 3, 4, 2, 1).reshape(-1, 3).sum(1) > 0
    assert set(torch.nonzero(moved)[:, 0].tolist()) <= set(base["src_mask"].tolist())
    assert 0.001 < float(diff.max()) < 0.05
    # rigid perturbation of either side: pose_new(src_new) == the same target-frame points as before
    for perturb_source in (True, False):
        d = pkg.augment.rigid_perturb(clone(base), 0.1, gen, perturb_source=perturb_source)
        P0, P1 = base["pose"][0], d["pose"][0]
        src0, src1 = masked_xyz(base, "src"), masked_xyz(d, "src")
        tgt0, tgt1 = masked_xyz(base, "tgt"), masked_xyz(d, "tgt")
        if perturb_source:
            a = src0 @ P0[:3, :3].T + P0[:3, 3]
            b = src1 @ P1[:3, :3].T + P1[:3, 3]
            assert torch.equal(tgt0, tgt1) and not torch.equal(src0, src1)
        else:      # the target moved by M: pose_new = M pose, so pose_new(src) = M (pose(src))
            M = P1 @ torch.linalg.inv(P0)
            a = (src0 @ P0[:3, :3].T + P0[:3, 3]) @ M[:3, :3].T + M[:3, 3]
            b = src1 @ P1[:3, :3].T + P1[:3, 3]
            assert torch.equal(src0, src1) and (tgt1 - (tgt0 @ M[:3, :3].T + M[:3, 3])).abs().max() < 1e-5
        assert (a - b).abs().max() < 1e-5
        rot = P1[:3, :3]
        assert (rot @ rot.T - torch.eye(3)).abs().max() < 1e-5
    # swap
    d = pkg.augment.random_swap(clone(base), swap=True)
    assert torch.equal(d["src_mask"], base["tgt_mask"]) and torch.equal(d["tgt_xyz_rgba"], base["src_xyz_rgba"])
    assert (d["pose"][0] @ base["pose"][0] - torch.eye(4)).abs().max() < 1e-5
    s = pkg.augment.sample_se3_small(0.1, gen)
    assert s.shape == (4, 4) and abs(float(torch.det(s[:3, :3])) - 1.0) < 1e-5


def test_occupancy_grid_fill_semantics(pkg):
    """nerfacc 0.3.5 OccupancyGrid.every_n_step restated (train_ngp_nerf.py:293): warm-up touches every cell, the
    EMA keeps the running maximum, the binary grid thresholds at min(mean, occ_thre)."""
    import torch
    og = pkg.OccupancyGrid([-1.5] * 3 + [1.5] * 3, 16).train()
    ball = lambda x: (x.norm(dim=1) < 1.0).float() * 0.5
    og.every_n_step(0, ball)
    n0 = int(og.binary.sum())
    assert og.binary.shape == (16, 16, 16) and 300 < n0 < 900          # ~ the cells of a radius-1 ball in a 3^3 box
    og.every_n_step(3, lambda x: torch.ones(x.shape[0]))                # not a multiple of n = 16: no update
    assert int(og.binary.sum()) == n0
    before = og.occs.clone()
    og.every_n_step(16, lambda x: torch.zeros(x.shape[0]))              # decay only
    assert torch.allclose(og.occs, before * 0.95)
    og.every_n_step(512, ball, generator=torch.Generator().manual_seed(0))   # past warm-up: a subset of the cells
    assert int(og.binary.sum()) > 0
    og.eval()
    import pytest
    with pytest.raises(RuntimeError):
        og.every_n_step(0, ball)


def test_pair_pipeline_host_logic(pkg):
    """pipeline.PairPipeline without a GPU: results in item order, one engine slot per worker thread, the first
    exception reaches the caller and the pipeline stays usable."""
    import threading
    import time
    from importlib import import_module
    import torch
    nr = import_module("dreg-nerf_b200.nerf_regtr")
    seen = set()

    def fn(i):
        seen.add((threading.current_thread().name, nr.engine_slot(), torch.is_grad_enabled()))
        time.sleep(0.005)
        return i * i

    with pkg.PairPipeline(torch.device("cpu"), streams=3) as pipe:
        with torch.no_grad():
            assert pipe.map(fn, range(10)) == [i * i for i in range(10)]
        assert {s for _, s, _ in seen} == {0, 1, 2} and not any(g for _, _, g in seen)   # grad mode follows the caller
        with pytest.raises(ZeroDivisionError):
            pipe.map(lambda i: 1 / (i - 3), range(8))
        assert pipe.map(fn, range(4)) == [0, 1, 4, 9]
    assert nr.engine_slot() == 0                                  # the caller's thread keeps slot 0
    assert pkg.PairPipeline(None, streams=1).map(fn, [2, 3]) == [4, 9]


def test_reference_arm_does_not_map_the_product_library():
    """bench.py --impl reference builds its inputs with the package's pure-Python pieces (module skeleton, seeded
    state dict, synthetic pairs, NGP field mirror) - none of them may dlopen libdregb200.so."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, gc; sys.path.insert(0, %r)\n"
        "import torch, dreg_nerf_b200 as pkg\n"
        "from oracle.make_goldens import make_field\n"
        "m = pkg.NeRFRegTr(); sd = pkg.synthetic.seeded_state_dict(m, seed=0, attn_gain=4.0); del m; gc.collect()\n"
        "pkg.synthetic.make_pair(res=32, pair_id=0); make_field(pkg, 500, 8.0); pkg.synthetic.extract_scene(32, 4)\n"
        "print('MAPPED' if 'libdregb200' in open('/proc/self/maps').read() else 'CLEAN')\n" % root)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.strip().endswith("CLEAN")
