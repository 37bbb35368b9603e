"""pipeline.PairPipeline: pairs of a batch in flight on several CUDA streams (one engine per slot, shared
parameters) give the results of the sequential loop."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _model(pkg, cuda, precision="fp32", training=True):
    torch.manual_seed(0)
    model = pkg.NeRFRegTr(precision=precision)
    model.load_state_dict(pkg.synthetic.seeded_state_dict(model, seed=0, attn_gain=4.0))
    model = model.to(cuda).train(training)
    model.correspondence_decoder.q_norm.requires_grad_(False)
    return model


def test_pipeline_forward_is_bit_identical_to_the_loop(pkg, cuda):
    """Every pair runs the same kernels in the same order on its own stream and engine: same bits."""
    model = _model(pkg, cuda)
    pairs = [pkg.synthetic.to_device(pkg.synthetic.make_pair(res=32, pair_id=i), cuda) for i in range(7)]
    keys = ("src_feats", "tgt_feats", "src_kp_warped", "tgt_kp_warped", "src_overlap")

    def fwd(data):
        out = model(dict(data))
        return [out[k][0] for k in keys] + [out["pose"]]

    with torch.no_grad():
        want = [fwd(p) for p in pairs]
        rv_before = model.fpn3d.backbone_net.bn1.running_var.clone()
        with pkg.PairPipeline(cuda, streams=3) as pipe:
            got = pipe.map(fwd, pairs)
            got2 = pipe.map(fwd, pairs[::-1])[::-1]          # another assignment of pairs to slots
        torch.cuda.synchronize()
    assert len(model._engines) == 3, "one engine per slot"
    for a, b, c in zip(want, got, got2):
        for x, y, z in zip(a, b, c):
            assert torch.equal(x, y) and torch.equal(x, z)
    # running statistics: updated by slot 0's pairs only - they did move, and stayed finite
    rv = model.fpn3d.backbone_net.bn1.running_var
    assert torch.isfinite(rv).all() and not torch.equal(rv, rv_before)


def test_pipeline_training_step_accumulates_the_same_gradients(pkg, cuda):
    """forward + loss + backward of 5 pairs through the pipeline: gradients equal to the sequential accumulation up
    to the order of the fp32 sums.  Running-statistics BatchNorm as in tests/golden/grad_32_eval.pt: with batch
    statistics the deepest stage of a 32^3 grid normalises ONE voxel per grid - its true gradient is zero and what
    is computed is amplified rounding noise, different even between two sequential runs."""
    from oracle.make_goldens import training_loss
    model = _model(pkg, cuda, precision="fp32", training=False)
    pairs = [pkg.synthetic.to_device(pkg.synthetic.make_pair(res=32, pair_id=10 + i), cuda) for i in range(5)]
    named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
    params = [p for _, p in named]

    def one(data):
        loss = training_loss(model(dict(data)))
        (loss / len(pairs)).backward()
        return loss.detach()

    def grads():
        return [p.grad.detach().clone() if p.grad is not None else None for p in params]

    model.zero_grad(set_to_none=True)
    want_loss = torch.stack([one(p) for p in pairs])
    want = grads()
    model.zero_grad(set_to_none=True)
    with pkg.PairPipeline(cuda, streams=3) as pipe:
        got_loss = torch.stack(pipe.map(one, pairs))
    torch.cuda.synchronize()
    got = grads()
    assert torch.equal(want_loss, got_loss), "per-pair losses are bit-identical"
    worst = 0.0
    largest = max(w.abs().max().item() for w in want if w is not None)
    for (name, _), w, g in zip(named, want, got):
        assert (w is None) == (g is None)
        if w is not None:
            # k_proj.bias: a constant added to all logits of a soft-max, its true gradient is zero
            scale = largest if name.endswith("k_proj.bias") else w.abs().max().clamp_min(1e-30).item()
            worst = max(worst, (w - g).abs().max().item() / scale)
    print("pipeline vs loop, accumulated gradients: worst rel diff %.2e" % worst)
    assert worst < 1e-5


def test_pipeline_propagates_errors(pkg, cuda):
    model = _model(pkg, cuda)
    good = pkg.synthetic.to_device(pkg.synthetic.make_pair(res=32, pair_id=0), cuda)
    bad = dict(good, src_mask=good["src_mask"].clone())
    bad["src_mask"][5] = 32 ** 3 + 17                                          # index outside the grid
    with pkg.PairPipeline(cuda, streams=2) as pipe, torch.no_grad():
        with pytest.raises(pkg.DrbError):
            pipe.map(lambda d: model(dict(d))["pose"], [good, bad, good, good])
        pose = pipe.map(lambda d: model(dict(d))["pose"], [good, good])       # the pipeline is still usable
    assert torch.equal(pose[0], pose[1])


def test_pipeline_full_path_is_bit_identical_to_the_loop(pkg, cuda):
    """extract (two NeRF blocks -> voxel grids) + register per pair, several pairs in flight: every kernel of the
    path - marcher, field queries, CUB sorts, engine - on a worker's own stream gives the loop's bits."""
    res, n_cam = 64, 8
    model = _model(pkg, cuda)
    occ, poses = pkg.synthetic.extract_scene(res, n_cam)
    meta = dict(pkg.synthetic.extract_meta(poses), camera_poses=poses.to(cuda))
    occ_d = occ.to(cuda)
    sgrid = pkg.SampleGrid(list(pkg.synthetic.AABB), res)
    k = int(occ.sum())
    fields = [[pkg.synthetic.make_ngp_field(seed=700 + 2 * i + s).to(cuda) for s in (0, 1)] for i in range(5)]
    gen = torch.Generator().manual_seed(5)
    jitters = [[torch.rand(k, 3, generator=gen).to(cuda) for _ in (0, 1)] for _ in range(5)]

    def one(i):
        grids = [pkg.extract_block(f, sgrid, occ_d, meta, cuda, jitter=j) for f, j in zip(fields[i], jitters[i])]
        if min(g[1].numel() for g in grids) < 16:
            return grids[0][0], grids[1][0], grids[0][1], grids[1][1]
        out = model({"src_xyz_rgba": grids[0][0].permute(3, 2, 0, 1).unsqueeze(0), "src_mask": grids[0][1],
                     "tgt_xyz_rgba": grids[1][0].permute(3, 2, 0, 1).unsqueeze(0), "tgt_mask": grids[1][1]})
        return grids[0][0], grids[1][0], grids[0][1], grids[1][1], out["pose"], out["src_feats"][0]

    with torch.no_grad():
        want = [one(i) for i in range(5)]
        with pkg.PairPipeline(cuda, streams=3) as pipe:
            got = pipe.map(one, range(5))
        torch.cuda.synchronize()
    kept = [int(w[2].numel()) for w in want]
    print("masked cells per source block:", kept)
    assert max(kept) > 0
    for a, b in zip(want, got):
        assert len(a) == len(b)
        for x, y in zip(a, b):
            assert torch.equal(x, y)
