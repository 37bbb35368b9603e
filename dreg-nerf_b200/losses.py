"""Host-side mirrors of the registration training losses either side of the hot path (SURVEY.md section 8f,
rank 1): ``CorrespondenceLoss`` (conerf/loss/correspondence_loss.py:7-51) and ``InfoNCELoss``
(conerf/loss/feature_loss.py:4-73), with the SE(3) helpers they use (conerf/register/se3.py:36-86).

Same constructor arguments, same ``forward`` signatures, same parameter name (``InfoNCELoss.W``) so that the
``feature_loss`` entry of a reference checkpoint (train_nerf_regtr.py:297) loads.  They consume the outputs of
``NeRFRegTr.forward`` and feed gradients back into ``drb_engine_backward`` through autograd.

``CorrespondenceLoss`` depends on ``robust_loss_pytorch.general.lossfun`` (Barron's general robust loss; the
package is not vendored in the reference and not installed here).  Its call site fixes ``alpha = 1``,
``scale = 0.5`` (correspondence_loss.py:30-34), for which the published loss is the Charbonnier /
pseudo-Huber form ``sqrt((x / c)^2 + 1) - 1``; restated here (parity with the wheel unpinned).
"""
from typing import List, Union

import torch
from torch import nn


def se3_inv(pose: torch.Tensor) -> torch.Tensor:
    """se3.py:36-41."""
    rot, trans = pose[..., :3, :3], pose[..., :3, 3:4]
    irot = rot.transpose(-1, -2)
    return torch.cat([irot, -irot @ trans], dim=-1)


def se3_transform_list(pose: Union[List[torch.Tensor], torch.Tensor], xyz: List[torch.Tensor]) -> list:
    """se3.py:63-86: xyz[b] [..., N, 3] moved by pose[b] [..., 3(4), 4]."""
    out = []
    for b in range(len(xyz)):
        assert xyz[b].shape[-1] == 3 and pose[b].shape[:-2] == xyz[b].shape[:-2]
        rot, trans = pose[b][..., :3, :3], pose[b][..., :3, 3:4]
        moved = rot @ xyz[b].transpose(-1, -2) + trans.expand(-1, xyz[b].shape[0])
        out.append(moved.transpose(-1, -2))
    return out


def robust_charbonnier(x: torch.Tensor, scale: float = 0.5) -> torch.Tensor:
    """robust_loss_pytorch.general.lossfun(x, alpha=1, scale) restated: (b / a) (((x / c)^2 / b + 1)^(a / 2) - 1)
    with a = 1, b = |a - 2| = 1."""
    return torch.sqrt(torch.square(x / scale) + 1.0) - 1.0


class CorrespondenceLoss(nn.Module):
    """correspondence_loss.py:7-51."""

    def __init__(self, metric: str = "mae", robust_loss: bool = True) -> None:
        super().__init__()
        assert metric in ["mse", "mae"]
        self.metric = metric
        self.robust_loss_func = robust_loss

    def forward(self, kp_before, kp_warped_pred, pose_gt, overlap_weights=None, eps=1e-6):
        kp_warped_gt = se3_transform_list(pose_gt, kp_before)
        corr_err = torch.cat(kp_warped_pred, dim=0) - torch.cat(kp_warped_gt, dim=0)
        if self.robust_loss_func:
            corr_err = robust_charbonnier(corr_err, 0.5)
        if self.metric == "mae":
            corr_err = torch.sum(torch.abs(corr_err), dim=-1)
        else:
            corr_err = torch.sum(torch.square(corr_err), dim=-1)
        if overlap_weights is not None:
            overlap_weights = torch.cat(overlap_weights)
            return torch.sum(overlap_weights * corr_err) / torch.clamp_min(torch.sum(overlap_weights), eps)
        return torch.mean(corr_err, dim=1)


class InfoNCELoss(nn.Module):
    """feature_loss.py:4-73 (positives: nearest point within r_p; points within r_n are ignored)."""

    def __init__(self, d_embed, r_p, r_n) -> None:
        super().__init__()
        self.r_p = r_p
        self.r_n = r_n
        self.n_sample = 256
        self.W = nn.Parameter(torch.zeros(d_embed, d_embed), requires_grad=True)
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.normal_(self.W, std=0.1)

    def compute_infonce(self, anchor_feat, positive_feat, anchor_xyz, positive_xyz):
        w_triu = torch.triu(self.W)
        w_sym = w_triu + w_triu.T
        match_logits = torch.einsum("...ic,cd,...jd->...ij", anchor_feat, w_sym, positive_feat)
        with torch.no_grad():
            dist = torch.cdist(anchor_xyz, positive_xyz)
            dist1, idx1 = dist.topk(k=1, dim=-1, largest=False)
            mask = dist1[..., 0] < self.r_p
            ignore = dist < self.r_n
            ignore.scatter_(-1, idx1, 0)
        match_logits[..., ignore] = -float("inf")
        loss = -torch.gather(match_logits, -1, idx1).squeeze(-1) + torch.logsumexp(match_logits, dim=-1)
        return torch.sum(loss[mask]) / torch.sum(mask)

    def forward(self, src_feat, tgt_feat, src_xyz, tgt_xyz):
        losses = [self.compute_infonce(src_feat[b], tgt_feat[b], src_xyz[b], tgt_xyz[b]) for b in range(len(src_feat))]
        return torch.mean(torch.stack(losses))


class RegistrationLoss(nn.Module):
    """The part of train_nerf_regtr.py:171-228's objective that does not re-enter a NeRF checkpoint:
    ``feature`` (weight 0.1, :109,204-208) + ``corr`` in both directions (weight 1, :110,212-224), with optional
    overlap targets (``compute_visibility_score``) as correspondence weights and for the BCE ``overlap`` term in the
    reference's own argument order (:195).  Used by bench.py's training stage with the synthetic ground-truth pose."""

    def __init__(self, d_embed=256, r_p=0.2, r_n=0.4):
        super().__init__()
        self.feature_loss = InfoNCELoss(d_embed, r_p, r_n)
        self.corr_loss = CorrespondenceLoss(metric="mae")
        self.overlap_loss = nn.BCEWithLogitsLoss()
        self.weight = {"overlap": 1.0, "feature": 0.1, "corr": 1.0}

    def forward(self, pred, pose_gt, src_overlap_gt=None, tgt_overlap_gt=None):
        pose34 = pose_gt[..., :3, :]
        losses = {}
        if src_overlap_gt is not None and tgt_overlap_gt is not None:
            gt = torch.cat(list(src_overlap_gt) + list(tgt_overlap_gt), dim=-2)
            pr = torch.cat(pred["src_overlap"] + pred["tgt_overlap"], dim=-2)
            losses["overlap"] = self.overlap_loss(gt[-1], pr[-1])
        losses["feature"] = self.feature_loss([f[-1] for f in pred["src_feats"]], [f[-1] for f in pred["tgt_feats"]],
                                              se3_transform_list(pose34, pred["src_kp"]), pred["tgt_kp"])
        # the reference always passes the visibility targets as weights, lists of [num_layers, N, 1] (:212-223); their
        # broadcast against the [N] error vector (correspondence_loss.py:46-48) makes the result sum_j err_j whenever
        # sum w > eps - reproduced as is.  Without targets the weights are ones.
        def weights(gt, warped):
            return list(gt) if gt is not None else [torch.ones_like(w[..., :1]) for w in warped]
        src = self.corr_loss(pred["src_kp"], [w[-1] for w in pred["src_kp_warped"]], pose34,
                             overlap_weights=weights(src_overlap_gt, pred["src_kp_warped"]))
        tgt = self.corr_loss(pred["tgt_kp"], [w[-1] for w in pred["tgt_kp_warped"]],
                             torch.stack([se3_inv(p) for p in pose34]),
                             overlap_weights=weights(tgt_overlap_gt, pred["tgt_kp_warped"]))
        losses["corr"] = src + tgt
        total = sum(losses[k] * self.weight[k] for k in losses)
        return total, losses
