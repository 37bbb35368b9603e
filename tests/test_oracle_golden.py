"""Pins the oracle (CPU, no GPU): against the fixtures produced from the REFERENCE's own modules by
oracle/make_goldens.py, against the reference itself when /root/reference is present, and against
the known-answer vectors of SURVEY.md section 4."""
import math
import os

import pytest
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _rel(a, b):
    return ((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30)).item()


@pytest.fixture(scope="module")
def holder(pkg):
    torch.manual_seed(0)
    return pkg.NeRFRegTr()


@pytest.mark.parametrize("case", ["fwd_32_eval", "fwd_64_train"])
def test_oracle_matches_reference_fixture(pkg, holder, case):
    """oracle/regtr.py reproduces the reference's outputs (bit-exact on the generating machine; a
    different CPU ISA may round differently, hence the 1e-4 guard instead of equality)."""
    from oracle import regtr
    fix = torch.load(os.path.join(GOLDEN, case + ".pt"))
    sd = pkg.synthetic.seeded_state_dict(holder, seed=fix["seed"], attn_gain=fix["gain"])
    data = pkg.synthetic.make_pair(res=fix["res"], pair_id=fix["pair_id"])
    assert [int(data["src_mask"].sum()), int(data["tgt_mask"].sum()), data["src_mask"].numel(),
            data["tgt_mask"].numel()] == fix["mask_digest"].tolist(), "synthetic inputs changed"
    with torch.no_grad():
        out = regtr.forward(sd, data, training=fix["train"])
    assert out["src_kp"][0].shape == fix["src_kp"].shape and out["tgt_kp"][0].shape == fix["tgt_kp"].shape
    for k in ("src_kp", "tgt_kp", "src_kp_warped", "tgt_kp_warped", "src_overlap", "tgt_overlap"):
        assert _rel(out[k][0], fix[k]) < 1e-4, k
    assert _rel(out["pose"], fix["pose"]) < 1e-4
    assert _rel(out["src_feats"][0][:, ::37, ::5], fix["src_feats_sample"]) < 1e-4
    assert _rel(out["tgt_feats"][0][:, ::37, ::5], fix["tgt_feats_sample"]) < 1e-4


def test_oracle_vs_live_reference(pkg, holder):
    """Where the reference tree exists (build container) compare against it directly, bit-exact."""
    from oracle import regtr
    from oracle.ref_shim import import_reference, reference_available
    if not reference_available():
        pytest.skip("reference tree not present on this machine")
    ref = import_reference()
    torch.manual_seed(0)
    rm = ref.NeRFRegTr()
    sd = pkg.synthetic.seeded_state_dict(rm, seed=1, attn_gain=2.0)
    rm.load_state_dict(sd)
    rm.eval()
    data = pkg.synthetic.make_pair(res=32, pair_id=3)
    with torch.no_grad():
        a = rm({k: (v.clone() if torch.is_tensor(v) else v) for k, v in data.items()})
        b = regtr.forward(sd, data, training=False)
    for k in ("src_feats", "tgt_feats", "src_kp_warped", "tgt_kp_warped", "src_overlap", "tgt_overlap"):
        assert torch.equal(a[k][0], b[k][0]), k
    assert torch.equal(a["pose"], b["pose"])


def test_known_answer_vectors():
    from oracle import regtr, extract
    kat = torch.load(os.path.join(GOLDEN, "kat.pt"))
    pe = regtr.pos_embed_sine(kat["pe_in"])
    assert _rel(pe, kat["pe_out"]) < 1e-6
    # SURVEY section 4 [measured]: PositionEmbeddingCoordsSine(3,256)([[0.1,0.2,0.3]])
    assert pe.shape == (2, 256) and float(pe[0, 252:].abs().sum()) == 0.0
    assert abs(float(pe[0].sum()) - 138.747589) < 1e-3
    assert torch.allclose(pe[0, :6], torch.tensor([0.587785, 0.809017, 0.508145, 0.861272, 0.436938, 0.899492]), atol=1e-5)
    assert _rel(regtr.compute_rigid_transform(kat["proc_a"], kat["proc_b"], kat["proc_w"]), kat["proc_out"]) < 1e-5
    # the single golden vector of the reference repo: conerf/utils/nerfacc_utils.py:56-63
    t = extract.transmittance_from_alpha(kat["alphas"], kat["ray_indices"])
    assert torch.allclose(t, kat["transmittance"], atol=1e-6)
    vis = (t >= 0.3) & (kat["alphas"] >= 0.2)
    assert vis.tolist() == [True, True, False, True, False, False, True]


def test_procrustes_recovers_known_transform():
    from oracle import regtr
    g = torch.Generator().manual_seed(0)
    a = torch.randn(1, 50, 3, generator=g, dtype=torch.float64)
    R = torch.tensor([[math.cos(0.7), -math.sin(0.7), 0], [math.sin(0.7), math.cos(0.7), 0], [0, 0, 1.0]],
                     dtype=torch.float64)
    t = torch.tensor([0.3, -0.2, 0.5], dtype=torch.float64)
    out = regtr.compute_rigid_transform(a, a @ R.t() + t, torch.rand(1, 50, generator=g, dtype=torch.float64))
    assert (out[0, :, :3] - R).abs().max() < 1e-8 and (out[0, :, 3] - t).abs().max() < 1e-8


def test_downsample_properties():
    """R4 restatement: averaging is idempotent per cell, lengths add up, early exit at <= 3000."""
    from oracle.downsample import batched_grid_subsample, hierarchical_grid_subsample, subsample_dl_schedule
    assert [round(v, 6) for v in subsample_dl_schedule(4)] == [0.05, 0.1, 0.2, 0.4]
    g = torch.Generator().manual_seed(2)
    pts = torch.rand(6000, 3, generator=g) * 2 - 1
    feats = torch.randn(6000, 8, generator=g)
    rows, lens = batched_grid_subsample(pts, feats, torch.tensor([3500, 2500]), 0.2)
    assert rows.shape[0] == int(lens.sum())
    again, lens2 = batched_grid_subsample(rows[:, :3], rows[:, 3:], lens, 0.2)
    assert lens2.tolist() == lens.tolist() and torch.allclose(again, rows, atol=1e-6)
    p, f, l = hierarchical_grid_subsample(pts, feats, torch.tensor([3500, 2500]), 6)
    assert p.shape[0] == int(l.sum()) <= 3000
    # empty cloud and single point edge cases
    rows, lens = batched_grid_subsample(pts[:1], feats[:1], torch.tensor([1, 0]), 0.05)
    assert lens.tolist() == [1, 0] and torch.equal(rows[0, :3], pts[0])


def test_hash_grid_level_table():
    from oracle import ngp
    levels, total = ngp.level_table()
    assert total == 6299960                       # 6 299 960 entries
    assert [r for _, r, _, _ in levels[:5]] == [16, 24, 34, 49, 71]
    assert [n for _, _, n, _ in levels[:5]] == [4096, 13824, 39304, 117656, 357912]
    assert all(n == 1 << 19 for _, _, n, _ in levels[5:])


def test_vectorized_surface_mask_equals_scalar_oracle(pkg):
    """The lock-step (timed-baseline) marcher and the scalar restatement agree on the extract fixture."""
    from oracle import extract, ngp
    from oracle.make_goldens import make_field, extract_scene
    fix = torch.load(os.path.join(GOLDEN, "extract_32.pt"))
    _, ref = make_field(pkg, fix["seed"], fix["table_std"])
    occ, cams = extract_scene(fix["res"], fix["n_cam"])
    roi = [-1.5] * 3 + [1.5] * 3
    pts = extract.sample_points(fix["sub"], fix["jitter"], fix["res"], roi)
    assert torch.equal(pts, fix["points"])
    dens_fn = lambda x: ngp.query_density(x, ref["aabb"], ref["table"], ref["w1"], ref["w2"])[0]
    m = extract.surface_mask_vectorized(pts, cams, occ, fix["res"], roi, roi, fix["step"], 0.5, dens_fn)
    assert torch.equal(m, fix["surface_mask"])
    d, feat = ngp.query_density(pts, ref["aabb"], ref["table"], ref["w1"], ref["w2"])
    assert torch.allclose(d, fix["density"], rtol=1e-5, atol=1e-6)
    assert torch.equal(d > 0.7, fix["density_mask"])


def test_visibility_fixture_lockstep_oracle(pkg):
    """The lock-step CPU marcher (bench.py's cpu_baseline for the extract half) reproduces the scalar
    oracle's visibility fixture wherever the decision is not within 1e-3 of the cut-off."""
    import os
    from oracle import extract, ngp
    from oracle.make_goldens import extract_scene, make_field
    fix = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "visibility_32.pt"))
    ref = make_field(pkg, fix["seed"], fix["table_std"])[1]
    occ, cams = extract_scene(fix["res"], fix["n_cam"])
    roi = [-1.5, -1.5, -1.5, 1.5, 1.5, 1.5]
    dens_fn = lambda x: ngp.query_density(x, ref["aabb"], ref["table"], ref["w1"], ref["w2"])[0]
    assert (dens_fn(fix["points"]) - fix["density"]).abs().max() <= 1e-5 * fix["density"].abs().max()
    got = extract.surface_mask_vectorized(fix["points"], cams, occ, fix["res"], roi, roi, fix["step"], 0.5, dens_fn)
    band = (fix["best"] - 0.5).abs() < 1e-3
    assert int(((got != fix["visible"]) & ~band).sum()) == 0


def test_c_oracle_matches_python_oracle(pkg):
    """oracle/extract_c.c (C / OpenMP restatement of A1 + A5, used for full-size checks and as the timed CPU
    baseline) against the pure-Python oracles: density to fp32 summation-order noise, the scalar marcher's
    fixtures bit for bit (extract_32) / off the 1e-3 band around the cut-off (visibility_32; the C code keeps
    the transmittance in fp32 like nerfacc, the Python oracle in doubles)."""
    import os
    from oracle import extract_c, ngp
    from oracle.make_goldens import extract_scene, make_field
    golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    _, ref = make_field(pkg, 0, 1.0)
    x = torch.rand(3000, 3, generator=torch.Generator().manual_seed(1)) * 3.2 - 1.6
    d_py = ngp.query_density(x, ref["aabb"], ref["table"], ref["w1"], ref["w2"])[0]
    d_c = extract_c.query_density(x, ref["aabb"], ref["table"], ref["w1"], ref["w2"])
    inside = d_py > 0
    assert bool(((d_c == 0) == (d_py == 0)).all())
    assert float(((d_c - d_py).abs()[inside] / d_py[inside]).max()) < 5e-6
    roi = [-1.5, -1.5, -1.5, 1.5, 1.5, 1.5]
    fix = torch.load(os.path.join(golden, "extract_32.pt"))
    ref = make_field(pkg, fix["seed"], fix["table_std"])[1]
    occ, cams = extract_scene(fix["res"], fix["n_cam"])
    m, _, n_samples = extract_c.surface_mask(fix["points"], cams, occ, fix["res"], roi, roi, fix["step"], 0.5, ref)
    assert torch.equal(m, fix["surface_mask"]) and n_samples > 0
    fix = torch.load(os.path.join(golden, "visibility_32.pt"))
    ref = make_field(pkg, fix["seed"], fix["table_std"])[1]
    occ, cams = extract_scene(fix["res"], fix["n_cam"])
    m, best, _ = extract_c.surface_mask(fix["points"], cams, occ, fix["res"], roi, roi, fix["step"], 0.5, ref)
    band = (fix["best"] - 0.5).abs() < 1e-3
    assert int(((m != fix["visible"]) & ~band).sum()) == 0
    # restricting the rays to flagged points / marching every camera does not change the flagged points' answer
    act = torch.zeros(m.numel(), dtype=torch.bool)
    act[::3] = True
    m2, _, _ = extract_c.surface_mask(fix["points"], cams, occ, fix["res"], roi, roi, fix["step"], 0.5, ref,
                                      active=act, all_rays=True)
    assert torch.equal(m2[act], m[act]) and not bool(m2[~act].any())


def test_oracle_gradients_match_reference(pkg):
    """Backward pin for the round that builds the backward kernels: autograd through the functional oracle
    reproduces the gradients the REFERENCE's own modules produced (tests/golden/grad_32_eval.pt, written by
    oracle/make_goldens.py gradient_case with /root/reference imported)."""
    import os
    from oracle import regtr
    from oracle.make_goldens import training_loss
    fix = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "grad_32_eval.pt"))
    torch.manual_seed(0)
    model = pkg.NeRFRegTr()
    sd = pkg.synthetic.seeded_state_dict(model, seed=fix["seed"], attn_gain=fix["gain"])
    leaf = {k: (v.clone().requires_grad_(True) if k in fix["digests"] else v) for k, v in sd.items()}
    data = pkg.synthetic.make_pair(res=fix["res"], pair_id=fix["pair_id"])
    loss = training_loss(regtr.forward(leaf, data, training=False))
    assert abs(float(loss) - float(fix["loss"])) < 1e-6 * abs(float(fix["loss"]))
    loss.backward()
    assert len(fix["digests"]) == 293
    for k, want in fix["digests"].items():
        g = leaf[k].grad.double()
        got = torch.stack([g.sum(), g.abs().sum(), (g * g).sum()])
        scale = float(want[1]) + 1e-30
        assert abs(float(got[0] - want[0])) < 1e-4 * scale, k
        assert abs(float(got[1] - want[1])) < 1e-4 * scale, k
        assert abs(float(got[2] - want[2])) < 1e-4 * (float(want[2]) + 1e-30), k
    for k, want in fix["samples"].items():
        got = leaf[k].grad.reshape(-1)[:64]
        assert float((got - want).abs().max()) <= 1e-4 * float(want.abs().max() + 1e-12), k
