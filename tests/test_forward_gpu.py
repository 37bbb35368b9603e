"""Full NeRFRegTr.forward on the GPU against the CPU oracle (oracle == reference bit-exact, see
tests/test_oracle_golden.py) on identical seeded inputs and weights."""
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = 1e-3   # BASELINE.json north_star: features / attention / SE(3) within 1e-3 rel fp32


def _run(pkg, cuda, res, training, pair_id=0, gain=4.0, precision="fp32"):
    from oracle import regtr
    torch.manual_seed(0)
    model = pkg.NeRFRegTr(precision=precision)
    sd = pkg.synthetic.seeded_state_dict(model, seed=0, attn_gain=gain)
    model.load_state_dict(sd)
    model = model.to(cuda).train(training)
    data = pkg.synthetic.make_pair(res=res, pair_id=pair_id)
    with torch.no_grad():
        out = model(pkg.synthetic.to_device(data, cuda))
        torch.cuda.synchronize()
        running = {}
        ref = regtr.forward(sd, data, training=training, running=running)
    return model, sd, data, out, ref, running


def _relerr(a, b):
    a, b = a.detach().double().cpu(), b.double()
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()


def _check(out, ref, tol=TOL):
    assert out["src_kp"][0].shape == ref["src_kp"][0].shape, "token counts differ"
    assert out["tgt_kp"][0].shape == ref["tgt_kp"][0].shape
    assert torch.equal(out["src_kp"][0].cpu(), ref["src_kp"][0]), "down-sampled key points must be bit exact"
    assert torch.equal(out["tgt_kp"][0].cpu(), ref["tgt_kp"][0])
    errs = {k: _relerr(out[k][0], ref[k][0]) for k in
            ("src_feats", "tgt_feats", "src_kp_warped", "tgt_kp_warped", "src_overlap", "tgt_overlap")}
    errs["pose"] = _relerr(out["pose"], ref["pose"])
    print("relative errors:", {k: "%.2e" % v for k, v in errs.items()})
    for k, v in errs.items():
        assert v < tol, "%s relative error %.3e exceeds %.1e" % (k, v, tol)
    return errs


def test_forward_32_eval(pkg, cuda):
    """BASELINE.json config 1 shape (32^3, running-statistics BatchNorm: at 32^3 the deepest
    feature map is 1^3, where torch refuses batch statistics)."""
    model, sd, data, out, ref, _ = _run(pkg, cuda, 32, training=False)
    assert out["pose"].shape == (6, 1, 3, 4)
    _check(out, ref)


def test_forward_64_train_bn(pkg, cuda):
    """Batch-statistics BatchNorm (the eval script never calls .eval()) plus the
    running-statistics side effect."""
    model, sd, data, out, ref, running = _run(pkg, cuda, 64, training=True)
    _check(out, ref)
    msd = model.state_dict()
    for prefix, (rm, rv) in list(running.items())[:8] + list(running.items())[-4:]:
        assert _relerr(msd[prefix + ".running_mean"], rm) < 1e-4
        assert _relerr(msd[prefix + ".running_var"], rv) < 1e-4
    assert int(msd["fpn3d.backbone_net.bn1.num_batches_tracked"]) == 2


def test_stage_taps_64(pkg, cuda):
    """Per-stage parity (backbone c1..c5, pyramid p1) to localise any drift."""
    from oracle import regtr
    torch.manual_seed(0)
    model = pkg.NeRFRegTr()
    sd = pkg.synthetic.seeded_state_dict(model, seed=0, attn_gain=4.0)
    model.load_state_dict(sd)
    model = model.to(cuda).train(True)
    model.sparse_fpn = False            # the taps compare whole feature maps
    data = pkg.synthetic.make_pair(res=64, pair_id=1)
    cap = {}
    with torch.no_grad():
        model(pkg.synthetic.to_device(data, cuda))
        regtr.forward(sd, data, training=True, capture=cap)
    for name in ("c1", "c2", "c3", "c4", "c5", "p5", "p4", "p3", "p2", "p1"):
        for which, side in ((0, "src"), (1, "tgt")):
            ref = cap["%s_%s" % (side, name)][0].permute(1, 2, 3, 0).reshape(-1)
            got = model.tap(name, which).cpu()
            err = _relerr(got, ref)
            print(name, side, "%.2e" % err)
            assert err < 2e-4, "%s/%s drifted: %.3e" % (name, side, err)


def test_sparse_fpn_equals_dense(pkg, cuda):
    """Output-sparse level-1 FPN convolutions: every output the path reads is unchanged (the backbone's
    split-K layers add atomically, so agreement is to rounding, not bit for bit)."""
    torch.manual_seed(0)
    model = pkg.NeRFRegTr()
    model.load_state_dict(pkg.synthetic.seeded_state_dict(model, seed=0, attn_gain=4.0))
    model = model.to(cuda).train(True)
    data = pkg.synthetic.to_device(pkg.synthetic.make_pair(res=64, pair_id=2), cuda)
    outs = []
    for sparse in (True, False):
        model.sparse_fpn = sparse
        with torch.no_grad():
            outs.append(model(dict(data)))
    for k in ("src_feats", "tgt_feats", "src_kp_warped", "tgt_kp_warped"):
        assert _relerr(outs[0][k][0], outs[1][k][0].cpu()) < 1e-4, k
    assert torch.equal(outs[0]["src_kp"][0], outs[1]["src_kp"][0])


def test_forward_bf16_mode_runs(pkg, cuda):
    """precision='bf16' (single bf16 product): looser agreement, same shapes."""
    model, sd, data, out, ref, _ = _run(pkg, cuda, 32, training=False, precision="bf16")
    assert out["pose"].shape == (6, 1, 3, 4)
    assert torch.isfinite(out["pose"]).all()
    assert _relerr(out["src_feats"][0], ref["src_feats"][0]) < 0.15


# bf16 operands (BASELINE configs[2]-[4]): one bf16 product per MMA, fp32 accumulation, fp32 BatchNorm / LayerNorm /
# soft-max / Procrustes.  A random-weight network is ill conditioned (53 BatchNorms over a few voxels, sharp soft-max
# over near-random logits): bf16 rounding moves the fp32 result by up to 1e-1 on the features and tens of degrees on
# the pose, and any two bf16 evaluations that differ in one rounding tie drift apart by the same amount.  Two gates:
#   * stage by stage (test_bf16_stage_taps_64): against the oracle evaluated with every tensor-core operand rounded
#     to bf16 at the places where the engine writes its planes (oracle.regtr.operand_rounding) the first stages must
#     agree tightly - a wrong rounding point, scale or plane would show there - and the distance may only grow as
#     ties accumulate;
#   * end to end (test_forward_bf16_parity_gate): the GPU path may not be farther from the fp32 oracle than
#     BF16_SLACK x the bf16-operand oracle itself is (it loses what bf16 arithmetic loses, no more).
BF16_SLACK = 2.0
BF16_FLOOR = 1e-2


def _pose_errors(pose, ref):
    """RRE (degrees) / RTE of the last decoder layer's pose against the oracle's (eval_nerf_regtr.py:46-65 metric)."""
    a, b = pose[-1, 0].double().cpu(), ref[-1, 0].double()
    cos = ((a[:, :3].T @ b[:, :3]).trace() - 1.0) / 2.0
    rre = torch.rad2deg(torch.acos(cos.clamp(-1.0, 1.0))).item()
    return rre, (a[:, 3] - b[:, 3]).norm().item()


def _all_errs(out, ref):
    errs = {"feats": max(_relerr(out["src_feats"][0], ref["src_feats"][0]), _relerr(out["tgt_feats"][0], ref["tgt_feats"][0])),
            "kp_warped": max(_relerr(out["src_kp_warped"][0], ref["src_kp_warped"][0]),
                             _relerr(out["tgt_kp_warped"][0], ref["tgt_kp_warped"][0])),
            "overlap": max(_relerr(out["src_overlap"][0], ref["src_overlap"][0]),
                           _relerr(out["tgt_overlap"][0], ref["tgt_overlap"][0])),
            "pose": _relerr(out["pose"], ref["pose"])}
    errs["rre_deg"], errs["rte"] = _pose_errors(out["pose"], ref["pose"])
    return errs


@pytest.mark.parametrize("res,training", [(32, False), (64, True), (128, True)])
def test_forward_bf16_parity_gate(pkg, cuda, res, training):
    """precision='bf16': key points bit exact (the token selection never sees a bf16 value); features, soft
    correspondences and overlap no farther from the fp32 oracle than BF16_SLACK x the bf16-operand oracle is.  The
    pose of a random-weight network is a Procrustes fit of near-random correspondences: reported, not gated."""
    from oracle import regtr
    model, sd, data, out, ref32, _ = _run(pkg, cuda, res, training=training, precision="bf16")
    with torch.no_grad(), regtr.operand_rounding(regtr.bf16_round):
        ref = regtr.forward(sd, data, training=training)
    assert torch.equal(out["src_kp"][0].cpu(), ref32["src_kp"][0]) and torch.equal(out["tgt_kp"][0].cpu(), ref32["tgt_kp"][0])
    ours, theirs, between = _all_errs(out, ref32), _all_errs(ref, ref32), _all_errs(out, ref)
    tag = "%d^3 (%s BN)" % (res, "batch" if training else "running")
    print("bf16 GPU vs fp32 oracle at %s:" % tag, {k: "%.2e" % v for k, v in ours.items()})
    print("bf16-operand oracle vs fp32 oracle at %s (what bf16 costs):" % tag, {k: "%.2e" % v for k, v in theirs.items()})
    print("bf16 GPU vs bf16-operand oracle at %s:" % tag, {k: "%.2e" % v for k, v in between.items()})
    for k in ("feats", "kp_warped", "overlap"):
        assert ours[k] < BF16_SLACK * theirs[k] + BF16_FLOOR, \
            "%s: bf16 path %.3e from the fp32 oracle, bf16 arithmetic alone %.3e" % (k, ours[k], theirs[k])
    assert all(torch.isfinite(t).all() for t in (out["pose"], out["src_feats"][0], out["src_kp_warped"][0]))


def test_bf16_stage_taps_64(pkg, cuda):
    """bf16 mode stage by stage against the bf16-operand oracle (dense FPN, batch-statistics BatchNorm)."""
    from oracle import regtr
    torch.manual_seed(0)
    model = pkg.NeRFRegTr(precision="bf16")
    sd = pkg.synthetic.seeded_state_dict(model, seed=0, attn_gain=4.0)
    model.load_state_dict(sd)
    model = model.to(cuda).train(True)
    model.sparse_fpn = False
    data = pkg.synthetic.make_pair(res=64, pair_id=1)
    cap, cap32 = {}, {}
    with torch.no_grad():
        model(pkg.synthetic.to_device(data, cuda))
        regtr.forward(sd, data, training=True, capture=cap32)
        with regtr.operand_rounding(regtr.bf16_round):
            regtr.forward(sd, data, training=True, capture=cap)
    # bound against the bf16-operand oracle per stage: tight while no rounding tie has flipped, then growing
    # measured on B200: c1 3e-7 (the stem computes exactly what bf16 arithmetic computes), c2 2e-3, c3 1.4e-2, c4 4e-2,
    # c5 2.5e-1 (BatchNorm over 8 voxels), p5..p1 1.4e-1 ... 3.7e-2 - always 2-3x closer than bf16 is to fp32
    bounds = {"c1": 1e-5, "c2": 1e-2, "c3": 5e-2, "c4": 1.5e-1, "c5": 6e-1, "p5": 4e-1, "p4": 3e-1, "p3": 1.5e-1,
              "p2": 1.5e-1, "p1": 1.5e-1}
    for name in ("c1", "c2", "c3", "c4", "c5", "p5", "p4", "p3", "p2", "p1"):
        for which, side in ((0, "src"), (1, "tgt")):
            ref = cap["%s_%s" % (side, name)][0].permute(1, 2, 3, 0).reshape(-1)
            ref32 = cap32["%s_%s" % (side, name)][0].permute(1, 2, 3, 0).reshape(-1)
            got = model.tap(name, which).cpu()
            err, cost = _relerr(got, ref), _relerr(ref, ref32)
            print("bf16 %s %s: vs bf16-operand oracle %.2e (bf16 itself costs %.2e vs fp32)" % (name, side, err, cost))
            assert err < bounds[name], "%s/%s: %.3e" % (name, side, err)


def test_256_cube_plumbing(pkg, cuda):
    """BASELINE.json configs[4] shape: a 256^3 block goes through extract and a 256^3 pair through
    NeRFRegTr.forward (index arithmetic beyond 2^24 cells, coarse occupancy bitmap with 8^3-voxel cells,
    engine buffers 8x the 128^3 ones).  Plumbing and finiteness; parity is covered at 32^3 / 64^3."""
    res = 256
    occ, poses = pkg.synthetic.extract_scene(res, 8)
    meta = dict(pkg.synthetic.extract_meta(poses), camera_poses=poses.to(cuda))
    sg = pkg.SampleGrid(list(pkg.synthetic.AABB), res)
    f = pkg.synthetic.make_ngp_field(seed=500).to(cuda)
    grid, mask = pkg.extract_block(f, sg, occ.to(cuda), meta, cuda)
    assert grid.shape == (res, res, res, 7) and mask.numel() > 1000
    rows = grid.reshape(-1, 7)
    assert bool((rows[mask].abs().sum(dim=1) > 0).all())
    assert int((rows.abs().sum(dim=1) > 0).sum()) == mask.numel()        # zeros outside the mask
    del grid, rows
    torch.manual_seed(0)
    model = pkg.NeRFRegTr().to(cuda).eval()
    data = pkg.synthetic.to_device(pkg.synthetic.make_pair(res=res, pair_id=0), cuda)
    with torch.no_grad():
        out = model(data)
    assert out["pose"].shape == (6, 1, 3, 4) and torch.isfinite(out["pose"]).all()
    assert all(torch.isfinite(t).all() for t in out["src_feats"])


def test_forward_128_train_bn(pkg, cuda):
    """BASELINE.json configs[1] size: the full 128^3 pair, batch-statistics BatchNorm (what the bench runs),
    against the CPU oracle (bit-identical to the reference's modules) - key points bit exact, features /
    correspondences / overlap / pose within the 1e-3 bar."""
    model, sd, data, out, ref, _ = _run(pkg, cuda, 128, training=True)
    assert out["pose"].shape == (6, 1, 3, 4)
    print("tokens", model.last_token_counts, "masked voxels", data["src_mask"].numel(), data["tgt_mask"].numel())
    _check(out, ref)


def test_stale_mask_is_rejected(pkg, cuda):
    """A mask built for another resolution must raise, not corrupt memory."""
    from importlib import import_module
    lib_mod = import_module("dreg-nerf_b200._lib")
    torch.manual_seed(0)
    model = pkg.NeRFRegTr().to(cuda).eval()
    good = pkg.synthetic.to_device(pkg.synthetic.make_pair(res=32, pair_id=0), cuda)
    bad = dict(good)
    bad["src_mask"] = good["src_mask"].clone()
    bad["src_mask"][5] = 32 ** 3 + 17
    with torch.no_grad():
        with pytest.raises(lib_mod.DrbError, match="mask index"):
            model(bad)
        out = model(dict(good))                     # the engine is still usable
    assert torch.isfinite(out["pose"]).all()
    bad["src_mask"][5] = -3
    with torch.no_grad(), pytest.raises(lib_mod.DrbError, match="mask index"):
        model(bad)
