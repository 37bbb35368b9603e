"""Generates tests/golden/*.pt from the REFERENCE's own modules (TEST INFRASTRUCTURE).

Run in the build container only (needs /root/reference):  python -m oracle.make_goldens
It (1) imports the reference through oracle/ref_shim.py, (2) loads the seeded synthetic state dict
into the reference's NeRFRegTr, (3) runs the reference forward on the seeded synthetic pairs,
(4) asserts that the functional oracle (oracle/regtr.py) reproduces it BIT-EXACTLY, and (5) writes
the reference's outputs as small fixtures that the CPU test-suite re-checks the oracle against on
any machine (the GPU box has no /root/reference).
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import regtr  # noqa: E402
from oracle.ref_shim import import_reference  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
CASES = [  # (name, resolution, batch-stat BN, pair id, decoder q/k gain)
    ("fwd_32_eval", 32, False, 0, 4.0),
    ("fwd_64_train", 64, True, 0, 4.0),
]


def _digest(t):
    t = t.double()
    return torch.stack([t.sum(), t.abs().sum(), (t * t).sum()])


def main():
    import dreg_nerf_b200 as pkg
    ref = import_reference()
    os.makedirs(GOLDEN, exist_ok=True)
    torch.manual_seed(0)
    model = ref.NeRFRegTr()
    # --- known-answer vectors of the importable reference pieces -----------------------------
    from conerf.register.position_embedding import PositionEmbeddingCoordsSine
    from conerf.register.se3 import compute_rigid_transform
    pe_in = torch.tensor([[0.1, 0.2, 0.3], [-1.2, 0.7, 1.45]])
    pe_out = PositionEmbeddingCoordsSine(3, 256, scale=1.0)(pe_in)
    g = torch.Generator().manual_seed(7)
    a = torch.randn(6, 50, 3, generator=g)
    b = torch.randn(6, 50, 3, generator=g)
    w = torch.rand(6, 50, generator=g)
    kat = {"pe_in": pe_in, "pe_out": pe_out, "proc_a": a, "proc_b": b, "proc_w": w,
           "proc_out": compute_rigid_transform(a, b, w),
           # conerf/utils/nerfacc_utils.py:56-63 (the only golden vector in the reference repo)
           "alphas": torch.tensor([0.4, 0.8, 0.1, 0.8, 0.1, 0.0, 0.9]),
           "ray_indices": torch.tensor([0, 0, 0, 1, 1, 2, 2]),
           "transmittance": torch.tensor([1.0, 0.6, 0.12, 1.0, 0.2, 1.0, 1.0])}
    assert torch.equal(regtr.pos_embed_sine(pe_in), pe_out)
    assert torch.allclose(regtr.compute_rigid_transform(a, b, w), kat["proc_out"], atol=1e-6)
    torch.save(kat, os.path.join(GOLDEN, "kat.pt"))
    keys = list(model.state_dict().keys())
    with open(os.path.join(GOLDEN, "state_dict_keys.txt"), "w") as fh:
        for k in keys:
            fh.write("%s %s\n" % (k, "x".join(str(s) for s in model.state_dict()[k].shape)))
    for name, res, train, pair_id, gain in CASES:
        sd = pkg.synthetic.seeded_state_dict(model, seed=0, attn_gain=gain)
        model.load_state_dict(sd)
        model.train(train)
        data = pkg.synthetic.make_pair(res=res, pair_id=pair_id)
        with torch.no_grad():
            out_ref = model({k: (v.clone() if torch.is_tensor(v) else v) for k, v in data.items()})
            out_or = regtr.forward(sd, data, training=train)
        for k in ("src_feats", "tgt_feats", "src_kp", "tgt_kp", "src_kp_warped", "tgt_kp_warped",
                  "src_overlap", "tgt_overlap"):
            assert torch.equal(out_ref[k][0], out_or[k][0]), (name, k)
        assert torch.equal(out_ref["pose"], out_or["pose"]), name
        fix = {"case": name, "res": res, "train": train, "pair_id": pair_id, "gain": gain, "seed": 0,
               "pose": out_ref["pose"],
               "src_kp": out_ref["src_kp"][0], "tgt_kp": out_ref["tgt_kp"][0],
               "src_kp_warped": out_ref["src_kp_warped"][0], "tgt_kp_warped": out_ref["tgt_kp_warped"][0],
               "src_overlap": out_ref["src_overlap"][0], "tgt_overlap": out_ref["tgt_overlap"][0],
               "src_feats_digest": _digest(out_ref["src_feats"][0]),
               "tgt_feats_digest": _digest(out_ref["tgt_feats"][0]),
               "src_feats_sample": out_ref["src_feats"][0][:, ::37, ::5].clone(),
               "tgt_feats_sample": out_ref["tgt_feats"][0][:, ::37, ::5].clone(),
               "mask_digest": torch.tensor([int(data["src_mask"].sum()), int(data["tgt_mask"].sum()),
                                            data["src_mask"].numel(), data["tgt_mask"].numel()]),
               "weights_digest": _digest(torch.cat([v.reshape(-1)[:1000] for v in sd.values()
                                                    if v.is_floating_point()]))}
        fix = {k: (v.clone().contiguous() if torch.is_tensor(v) else v) for k, v in fix.items()}
        torch.save(fix, os.path.join(GOLDEN, name + ".pt"))
        print("wrote", name, "tokens", fix["src_kp"].shape[0], fix["tgt_kp"].shape[0],
              "oracle == reference bit-exact")


def training_loss(out):
    """A fixed scalar of every output of NeRFRegTr.forward (stand-in for the training losses of
    train_nerf_regtr.py:171-256, which need NeRF checkpoints): used only to pin gradients."""
    loss = (out["pose"][-1] ** 2).mean()
    for side in ("src", "tgt"):
        loss = loss + (out[side + "_feats"][0][-1] ** 2).mean() + (out[side + "_kp_warped"][0] ** 2).mean() \
            + out[side + "_overlap"][0].mean()
    return loss


def gradient_case():
    """Backward fixture for the round that builds the backward kernels: gradients of ``training_loss`` w.r.t.
    every parameter, computed by autograd through the REFERENCE's own modules (32^3, running-statistics
    BatchNorm), asserted equal for the functional oracle, stored as per-parameter digests + a few samples."""
    import dreg_nerf_b200 as pkg
    ref = import_reference()
    torch.manual_seed(0)
    model = ref.NeRFRegTr()
    sd = pkg.synthetic.seeded_state_dict(model, seed=0, attn_gain=4.0)
    model.load_state_dict(sd)
    model.train(False)
    data = pkg.synthetic.make_pair(res=32, pair_id=0)
    out = model({k: (v.clone() if torch.is_tensor(v) else v) for k, v in data.items()})
    loss = training_loss(out)
    loss.backward()
    grads = {k: p.grad for k, p in model.named_parameters() if p.grad is not None}
    # the functional oracle on leaf copies of the same weights
    leaf = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and k in grads else v) for k, v in sd.items()}
    out_or = regtr.forward(leaf, data, training=False)
    loss_or = training_loss(out_or)
    loss_or.backward()
    assert torch.equal(loss_or.detach(), loss.detach()), (float(loss_or), float(loss))
    worst = 0.0
    for k, g in grads.items():
        go = leaf[k].grad
        assert go is not None, k
        worst = max(worst, float((go - g).abs().max() / (g.abs().max() + 1e-30)))
    assert worst < 1e-5, worst
    fix = {"loss": loss.detach(), "res": 32, "pair_id": 0, "gain": 4.0, "seed": 0,
           "digests": {k: _digest(g) for k, g in grads.items()},
           "samples": {k: grads[k].reshape(-1)[:64].clone() for k in list(grads)[:3] + list(grads)[-6:]},
           "oracle_vs_reference_max_rel": worst}
    torch.save(fix, os.path.join(GOLDEN, "grad_32_eval.pt"))
    print("wrote grad_32_eval: loss %.6f, %d parameter gradients, oracle vs reference max rel %.2e"
          % (float(loss), len(grads), worst))


def losses_case():
    """InfoNCELoss fixture from the REFERENCE's own module (conerf/loss/feature_loss.py is importable as is):
    inputs, loss value and gradients w.r.t. the features and W.  CorrespondenceLoss cannot be imported
    (robust_loss_pytorch is absent): its fixture is evaluated with the published closed form of Barron's loss at
    the call site's alpha = 1, scale = 0.5 in float64 - parity with the wheel itself is unpinned."""
    import_reference()
    from conerf.loss.feature_loss import InfoNCELoss
    from conerf.register.se3 import se3_transform_list, se3_inv
    g = torch.Generator().manual_seed(21)
    ref = InfoNCELoss(256, 0.2, 0.4)
    with torch.no_grad():
        ref.W.copy_(torch.randn(256, 256, generator=g) * 0.1)
    sf = torch.randn(300, 256, generator=g).requires_grad_(True)
    tf = torch.randn(280, 256, generator=g).requires_grad_(True)
    sx, tx = torch.rand(300, 3, generator=g), torch.rand(280, 3, generator=g)
    loss = ref([sf], [tf], [sx], [tx])
    gs, gt, gw = torch.autograd.grad(loss, [sf, tf, ref.W])
    pose = torch.eye(4)[:3][None].clone()
    pose[0, :, 3] = torch.tensor([0.1, -0.2, 0.3])
    kp, pred = torch.rand(50, 3, generator=g), torch.rand(50, 3, generator=g)
    w = torch.rand(6, 50, 1, generator=g)
    gt_pts = se3_transform_list(pose, [kp])[0].double()
    err = (torch.sqrt(((pred.double() - gt_pts) / 0.5) ** 2 + 1.0) - 1.0).abs().sum(-1)
    corr = (w.double() * err).sum() / w.double().sum().clamp_min(1e-6)
    fix = {"W": ref.W.detach().clone(), "sf": sf.detach(), "tf": tf.detach(), "sx": sx, "tx": tx,
           "infonce": loss.detach(), "g_sf": gs, "g_tf": gt, "g_W": gw,
           "pose": pose, "kp": kp, "pred": pred, "w": w, "corr": corr.float(), "pose_inv": se3_inv(pose)}
    torch.save(fix, os.path.join(GOLDEN, "losses.pt"))
    print("wrote losses: infonce %.6f corr %.6f" % (float(loss), float(corr)))


def extract_case():
    """Extract fixture: a 32^3 block of a seeded random-weight field, evaluated by the (slow, pure
    Python) oracle here so that the GPU box only has to load the result."""
    import math
    import dreg_nerf_b200 as pkg
    from oracle import extract
    res, n_pts, n_cam, table_std, step = 32, 120, 3, 3.0, 0.15
    field = make_field(pkg, seed=3, table_std=table_std)[1]
    occ, cams = extract_scene(res, n_cam)
    gen = torch.Generator().manual_seed(5)
    idx = torch.nonzero(occ.flatten())[:, 0]
    sub = idx[torch.randperm(idx.numel(), generator=gen)[:n_pts]].sort().values
    jitter = torch.rand(sub.numel(), 3, generator=gen)
    roi = [-1.5, -1.5, -1.5, 1.5, 1.5, 1.5]
    o = extract.extract_block(field, sub, jitter, occ, res, roi, roi, cams, step, use_c_marcher=True)
    o_py = extract.extract_block(field, sub, jitter, occ, res, roi, roi, cams, step)      # scalar Python marcher (doubles)
    print("extract_32: C marcher vs scalar Python marcher: %d differing surface decisions"
          % int((o["surface_mask"] != o_py["surface_mask"]).sum()))
    fix = {"res": res, "seed": 3, "table_std": table_std, "step": step, "n_cam": n_cam, "sub": sub, "jitter": jitter,
           "points": o["points"], "rgb": o["rgb"], "alpha": o["alpha"], "density": o["density"],
           "density_mask": o["density_mask"], "surface_mask": o["surface_mask"], "surface_best": o["surface_best"],
           "mask": o["mask"], "marcher": "oracle/extract_c.c (fp32, march_math.h)",
           "rows": o["grid"].reshape(-1, 7)[sub]}
    fix = {k: (v.clone().contiguous() if torch.is_tensor(v) else v) for k, v in fix.items()}
    torch.save(fix, os.path.join(GOLDEN, "extract_32.pt"))
    print("wrote extract_32: density>0.7 %d / %d, surface %d, both %d" % (
        int(o["density_mask"].sum()), sub.numel(), int(o["surface_mask"].sum()), o["mask"].numel()))


def visibility_case():
    """compute_visibility_score fixture (conerf/loss/confidence_loss.py:56-160): arbitrary query points -
    on and around the occupied shell, some outside the AABB - scored by the scalar oracle marcher, with
    the best surface-field value per point so that decisions within 1e-3 of the cut-off can be set aside."""
    import math
    import dreg_nerf_b200 as pkg
    from oracle import extract, ngp
    res, n_cam, seed, table_std = 32, 4, 7, 8.0
    step = 3.0 * math.sqrt(3) / 256
    ref = make_field(pkg, seed, table_std)[1]
    occ, cams = extract_scene(res, n_cam)
    roi = [-1.5, -1.5, -1.5, 1.5, 1.5, 1.5]
    g = torch.Generator().manual_seed(5)
    idx = torch.nonzero(occ.flatten())[:, 0]
    pick = idx[torch.randperm(idx.numel(), generator=g)[:360]]
    cell = torch.stack([pick // (res * res), (pick // res) % res, pick % res], dim=1).float()
    pts = (cell + torch.rand(360, 3, generator=g)) / res * 3.0 - 1.5
    pts = torch.cat([pts, torch.rand(40, 3, generator=g) * 3.6 - 1.8])
    dens_fn = lambda x: ngp.query_density(x, ref["aabb"], ref["table"], ref["w1"], ref["w2"])[0]
    from oracle import extract_c
    want, best, _ = extract_c.surface_mask(pts, cams, occ, res, roi, roi, step, 0.5, ref, all_rays=True)
    want_py, best_py = extract.surface_mask(pts, cams, occ, res, roi, roi, step, 0.5, dens_fn, return_best=True)
    print("visibility_32: C marcher vs scalar Python marcher: %d differing decisions, max |best| difference %.2e"
          % (int((want != want_py).sum()), float((best - best_py.float()).abs().max())))
    fix = {"res": res, "n_cam": n_cam, "seed": seed, "table_std": table_std, "step": step, "points": pts,
           "visible": want, "best": best.float(), "density": dens_fn(pts), "marcher": "oracle/extract_c.c (fp32, march_math.h)"}
    torch.save(fix, os.path.join(GOLDEN, "visibility_32.pt"))
    print("wrote visibility_32: %d of %d visible, %d within 1e-3 of the cut-off"
          % (int(want.sum()), pts.shape[0], int(((best - 0.5).abs() < 1e-3).sum())))


def make_field(pkg, seed, table_std):
    """(module, oracle parameter dict) of a seeded random-weight NGP field."""
    from importlib import import_module
    m = import_module("dreg-nerf_b200.ngp")
    f = pkg.synthetic.make_ngp_field(seed, table_std)
    p, c = f.mlp_base.params.detach().clone(), f.color_mlp.params.detach().clone()
    ref = {"aabb": f.aabb.clone(),
           "w1": p[:m.N_W1].reshape(64, 32), "w2": p[m.N_W1:m.N_W1 + m.N_W2].reshape(16, 64),
           "table": p[m.N_W1 + m.N_W2:].reshape(-1, 2),
           "c1": c[:m.N_C1].reshape(64, 32), "c2": c[m.N_C1:m.N_C1 + m.N_C2].reshape(64, 64),
           "c3": c[m.N_C1 + m.N_C2:].reshape(16, 64)}
    return f, ref


def extract_scene(res, n_cam):
    """(occupancy, camera centres) of the synthetic extract scene."""
    import dreg_nerf_b200 as pkg
    occ, poses = pkg.synthetic.extract_scene(res, n_cam)
    return occ, poses[:, :3, 3].contiguous()


if __name__ == "__main__":
    if "--gradient-only" in sys.argv:
        gradient_case()
        sys.exit(0)
    if "--losses-only" in sys.argv:
        losses_case()
        sys.exit(0)
    if "--visibility-only" in sys.argv:
        visibility_case()
        sys.exit(0)
    if "--extract-only" not in sys.argv:
        main()
    extract_case()
    visibility_case()
    gradient_case()
    losses_case()
