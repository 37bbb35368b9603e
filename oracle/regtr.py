"""R1-R3, R5-R10: functional CPU restatement of NeRFRegTr.forward (TEST INFRASTRUCTURE).

State-dict driven (the 772-key layout of the reference module), torch CPU ops as
the arithmetic library.  PINNED: ``oracle/make_goldens.py`` checks it against the
reference's own modules (imported through ``oracle/ref_shim.py``) and commits the
reference's outputs under ``tests/golden/``.

Reference files followed:
  conerf/register/nerf_regtr.py:112-248, 273-308, 350-394   (forward, decoder)
  conerf/model/resnet3d.py:76-172, conerf/model/feature_pyramid_net.py:39-108
  conerf/register/transformer.py:50-86, 225-299
  conerf/register/position_embedding.py:30-53
  conerf/register/se3.py:89-140
"""
import math

import torch
import torch.nn.functional as F

from oracle.downsample import hierarchical_grid_subsample

BN_EPS = 1e-5
BN_MOMENTUM = 0.1
LAYERS = (3, 4, 6, 3)            # resnet50, resnet3d.py:197-205
PLANES = (64, 128, 256, 512)


# Optional operand rounding (TEST INFRASTRUCTURE for the bf16 parity gate): inside `with operand_rounding(fn):` every
# tensor-core operand of the GPU path - inputs and weights of each convolution / linear layer, Q' K V P of the
# attention core, q / k of the decoder - is passed through `fn` (e.g. round to bf16) before it is used, at exactly the
# places where the engine writes its 16-bit planes.  Accumulation, bias, residuals, BatchNorm / LayerNorm / soft-max
# statistics and the Procrustes solve stay fp32.  Outside the context (`_R is None`) nothing changes: the functions
# below are the reference-pinned fp32 restatement.
_R = None


class operand_rounding:
    def __init__(self, fn):
        self.fn = fn

    def __enter__(self):
        global _R
        self.prev, _R = _R, self.fn
        return self

    def __exit__(self, *exc):
        global _R
        _R = self.prev
        return False


def bf16_round(t):
    return t.to(torch.bfloat16).to(torch.float32)


def _r(t):
    return t if _R is None else _R(t)


def _bn(x, sd, prefix, training, running=None):
    """nn.BatchNorm3d: batch statistics when ``training`` (the eval script never
    calls .eval(), eval_nerf_regtr.py:212-218), running statistics otherwise."""
    w, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    rm, rv = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    if training and running is not None:
        rm, rv = running.setdefault(prefix, (rm.clone(), rv.clone()))
        return F.batch_norm(x, rm, rv, w, b, True, BN_MOMENTUM, BN_EPS)
    if training:
        return F.batch_norm(x, None, None, w, b, True, BN_MOMENTUM, BN_EPS)
    return F.batch_norm(x, rm, rv, w, b, False, BN_MOMENTUM, BN_EPS)


def _bottleneck(x, sd, p, stride, has_down, training, running):
    """resnet3d.py:76-113."""
    out = F.relu(_bn(F.conv3d(_r(x), _r(sd[p + ".conv1.weight"])), sd, p + ".bn1", training, running))
    out = F.conv3d(_r(out), _r(sd[p + ".conv2.weight"]), stride=stride, padding=1)
    out = F.relu(_bn(out, sd, p + ".bn2", training, running))
    out = _bn(F.conv3d(_r(out), _r(sd[p + ".conv3.weight"])), sd, p + ".bn3", training, running)
    res = x
    if has_down:
        res = _bn(F.conv3d(_r(x), _r(sd[p + ".downsample.0.weight"]), stride=stride),
                  sd, p + ".downsample.1", training, running)
    return F.relu(out + res)


def resnet3d(x, sd, prefix="fpn3d.backbone_net", training=True, running=None, capture=None):
    """resnet3d.py:157-172 -> (c1..c5)."""
    c1 = F.conv3d(_r(x), _r(sd[prefix + ".conv1.weight"]), stride=2, padding=2)
    c1 = F.relu(_bn(c1, sd, prefix + ".bn1", training, running))
    feats = [c1]
    y = F.max_pool3d(c1, kernel_size=3, stride=2, padding=1)
    for li, (nblk, planes) in enumerate(zip(LAYERS, PLANES)):
        for bi in range(nblk):
            stride = 2 if (li > 0 and bi == 0) else 1
            y = _bottleneck(y, sd, "%s.layer%d.%d" % (prefix, li + 1, bi), stride, bi == 0,
                            training, running)
        feats.append(y)
    if capture is not None:
        for i, f in enumerate(feats):
            capture["c%d" % (i + 1)] = f
    return feats


def fpn3d(x, sd, training=True, running=None, capture=None):
    """feature_pyramid_net.py:63-105 (FeaturePyramid_v1) -> p1 [B,256,R/2,R/2,R/2]."""
    c1, c2, c3, c4, c5 = resnet3d(x, sd, training=training, running=running, capture=capture)
    fp = "fpn3d.feature_pyramid."

    def lateral(i, c, pad):
        return F.conv3d(_r(c), _r(sd[fp + "pyramid_transformation_%d.weight" % i]),
                        sd[fp + "pyramid_transformation_%d.bias" % i], padding=pad)

    def merge(i, top, lat):
        d, h, w = lat.shape[2:]
        up = F.interpolate(top, scale_factor=2)[:, :, :d, :h, :w]      # nearest, :58-61
        return F.conv3d(_r(up + lat), _r(sd[fp + "upsample_transform_%d.weight" % i]),
                        sd[fp + "upsample_transform_%d.bias" % i], padding=1)

    p5 = lateral(5, c5, 0)
    p4 = merge(4, p5, lateral(4, c4, 0))
    p3 = merge(3, p4, lateral(3, c3, 0))
    p2 = merge(2, p3, lateral(2, c2, 0))
    p1 = merge(1, p2, lateral(1, c1, 1))
    if capture is not None:
        capture.update(p5=p5, p4=p4, p3=p3, p2=p2, p1=p1)
    return p1


def pos_embed_sine(xyz, d_model=256, temperature=1000.0, scale=1.0):
    """position_embedding.py:30-53."""
    n_dim = xyz.shape[-1]
    npf = d_model // n_dim // 2 * 2
    pad = d_model - npf * n_dim
    dim_t = torch.arange(npf, dtype=torch.float32)
    dim_t = temperature ** (2 * torch.div(dim_t, 2, rounding_mode="trunc") / npf)
    v = (xyz * (scale * 2 * math.pi)).unsqueeze(-1) / dim_t
    emb = torch.stack([v[..., 0::2].sin(), v[..., 1::2].cos()], dim=-1)
    emb = emb.reshape(*xyz.shape[:-1], -1)
    return F.pad(emb, (0, pad))


def _mha(q_in, k_in, v_in, sd, p, nhead=8):
    """nn.MultiheadAttention forward (batch 1, no masks, dropout 0); q_in [Nq,D]."""
    d = q_in.shape[-1]
    w, b = sd[p + ".in_proj_weight"], sd[p + ".in_proj_bias"]
    q = F.linear(_r(q_in), _r(w[:d]), b[:d])
    k = F.linear(_r(k_in), _r(w[d:2 * d]), b[d:2 * d])
    v = F.linear(_r(v_in), _r(w[2 * d:]), b[2 * d:])
    hd = d // nhead
    if _R is None:
        qh = q.view(-1, nhead, hd).transpose(0, 1) * (1.0 / math.sqrt(hd))
        kh = k.view(-1, nhead, hd).transpose(0, 1)
        vh = v.view(-1, nhead, hd).transpose(0, 1)
        att = torch.softmax(qh @ kh.transpose(1, 2), dim=-1)
        o = (att @ vh).transpose(0, 1).reshape(-1, d)
    else:
        # the tensor-core attention kernel (csrc/attention.cu): Q' = q scale log2(e), K, V and P = 2^(S - max) are
        # 16-bit operands, the row sum is taken over the unrounded P
        qh = _r(q * (torch.tensor(1.0 / math.sqrt(hd), dtype=torch.float32) * 1.4426950408889634)).view(-1, nhead, hd).transpose(0, 1)
        kh = _r(k).view(-1, nhead, hd).transpose(0, 1)
        vh = _r(v).view(-1, nhead, hd).transpose(0, 1)
        sc = qh @ kh.transpose(1, 2)
        pr = torch.exp2(sc - sc.max(dim=-1, keepdim=True).values)
        o = ((_r(pr) @ vh) / pr.sum(dim=-1, keepdim=True)).transpose(0, 1).reshape(-1, d)
    return F.linear(_r(o), _r(sd[p + ".out_proj.weight"]), sd[p + ".out_proj.bias"])


def _ln(x, sd, p):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], 1e-5)


def cross_encoder(src, tgt, src_pos, tgt_pos, sd, prefix="transformer_encoder", num_layers=6):
    """transformer.py:50-86 + forward_pre :225-299; returns 2 x [L,N,D]."""
    outs_s, outs_t = [], []
    for li in range(num_layers):
        p = "%s.layers.%d" % (prefix, li)
        s2 = _ln(src, sd, p + ".norm1") + src_pos
        src = src + _mha(s2, s2, s2, sd, p + ".self_attn")
        t2 = _ln(tgt, sd, p + ".norm1") + tgt_pos
        tgt = tgt + _mha(t2, t2, t2, sd, p + ".self_attn")
        s2 = _ln(src, sd, p + ".norm2") + src_pos
        t2 = _ln(tgt, sd, p + ".norm2") + tgt_pos
        s3 = _mha(s2, t2, t2, sd, p + ".cross_attn")
        t3 = _mha(t2, s2, s2, sd, p + ".cross_attn")
        src, tgt = src + s3, tgt + t3
        s2 = _ln(src, sd, p + ".norm3")
        src = src + F.linear(_r(F.relu(F.linear(_r(s2), _r(sd[p + ".linear1.weight"]), sd[p + ".linear1.bias"]))),
                             _r(sd[p + ".linear2.weight"]), sd[p + ".linear2.bias"])
        t2 = _ln(tgt, sd, p + ".norm3")
        tgt = tgt + F.linear(_r(F.relu(F.linear(_r(t2), _r(sd[p + ".linear1.weight"]), sd[p + ".linear1.bias"]))),
                             _r(sd[p + ".linear2.weight"]), sd[p + ".linear2.bias"])
        outs_s.append(_ln(src, sd, prefix + ".norm"))
        outs_t.append(_ln(tgt, sd, prefix + ".norm"))
    return torch.stack(outs_s), torch.stack(outs_t)


def correspondence_decoder(src_f, tgt_f, src_xyz, tgt_xyz, sd, pos_scale=1.0,
                           prefix="correspondence_decoder"):
    """nerf_regtr.py:273-308,350-394.  src_f [L,Ns,D]."""
    d = src_f.shape[-1]
    s2 = src_f + pos_embed_sine(src_xyz, d, scale=pos_scale)
    t2 = tgt_f + pos_embed_sine(tgt_xyz, d, scale=pos_scale)

    def attend(qf, kf, val):
        q = _r(F.linear(_r(qf), _r(sd[prefix + ".q_proj.weight"]), sd[prefix + ".q_proj.bias"]) / math.sqrt(d))
        k = _r(F.linear(_r(kf), _r(sd[prefix + ".k_proj.weight"]), sd[prefix + ".k_proj.bias"]))
        att = torch.softmax(q @ k.transpose(-1, -2), dim=-1)
        return att @ val

    src_corr = attend(s2, t2, tgt_xyz)
    tgt_corr = attend(t2, s2, src_xyz)
    w, b = sd[prefix + ".conf_logits_decoder.weight"], sd[prefix + ".conf_logits_decoder.bias"]
    return src_corr, tgt_corr, torch.sigmoid(F.linear(src_f, w, b)), torch.sigmoid(F.linear(tgt_f, w, b))


def compute_rigid_transform(a, b, weights, eps=1e-6):
    """se3.py:89-140 (weighted branch)."""
    wn = weights[..., None] / torch.clamp_min(weights.sum(-1, keepdim=True)[..., None], eps)
    ca, cb = (a * wn).sum(-2), (b * wn).sum(-2)
    ac, bc = a - ca[..., None, :], b - cb[..., None, :]
    cov = ac.transpose(-2, -1) @ (bc * wn)
    u, _, vh = torch.linalg.svd(cov)
    v = vh.transpose(-1, -2)
    rot_pos = v @ u.transpose(-1, -2)
    v_neg = v.clone()
    v_neg[..., 2] *= -1
    rot_neg = v_neg @ u.transpose(-1, -2)
    rot = torch.where(torch.det(rot_pos)[..., None, None] > 0, rot_pos, rot_neg)
    t = -rot @ ca[..., :, None] + cb[..., :, None]
    return torch.cat((rot, t), dim=-1)


def gather_masked(grid_xyz_rgba, p1, mask):
    """nerf_regtr.py:138-147: trilinear (align_corners) upsample of p1 to the grid
    resolution, then rows at the flat (X,Y,Z) C-order indices ``mask``."""
    res = grid_xyz_rgba.shape[-3:]
    up = F.interpolate(p1, size=res, mode="trilinear", align_corners=True)
    c = up.shape[1]
    xyz = grid_xyz_rgba[:, :3].permute(0, 3, 4, 2, 1).reshape(1, -1, 3)[0, mask]
    feats = up.permute(0, 3, 4, 2, 1).reshape(1, -1, c)[0, mask]
    return xyz, feats


def forward(sd, data, num_downsample=6, pos_scale=1.0, training=True, running=None, capture=None):
    """NeRFRegTr.forward for one pair; returns the R10 dict."""
    src, tgt = data["src_xyz_rgba"], data["tgt_xyz_rgba"]
    if src.dim() == 6:
        src, tgt = src.squeeze(0), tgt.squeeze(0)
    src_mask, tgt_mask = data["src_mask"].reshape(-1), data["tgt_mask"].reshape(-1)
    cap_s = {} if capture is not None else None
    cap_t = {} if capture is not None else None
    p1_s = fpn3d(src[:, 3:], sd, training, running, cap_s)
    p1_t = fpn3d(tgt[:, 3:], sd, training, running, cap_t)
    src_xyz, src_feats = gather_masked(src, p1_s, src_mask)
    tgt_xyz, tgt_feats = gather_masked(tgt, p1_t, tgt_mask)
    if capture is not None:
        capture.update({"src_" + k: v for k, v in cap_s.items()})
        capture.update({"tgt_" + k: v for k, v in cap_t.items()})
        capture.update(src_gather_xyz=src_xyz, src_gather_feats=src_feats,
                       tgt_gather_xyz=tgt_xyz, tgt_gather_feats=tgt_feats)
    lengths = torch.tensor([src_xyz.shape[0], tgt_xyz.shape[0]], dtype=torch.int64)
    pts, feats, ds_len = hierarchical_grid_subsample(
        torch.cat([src_xyz, tgt_xyz]), torch.cat([src_feats, tgt_feats]), lengths, num_downsample)
    ns = int(ds_len[0])
    src_xyz, tgt_xyz = pts[:ns], pts[ns:]
    src_feats, tgt_feats = feats[:ns], feats[ns:]
    if capture is not None:
        capture.update(src_ds_xyz=src_xyz, tgt_ds_xyz=tgt_xyz, src_ds_feats=src_feats, tgt_ds_feats=tgt_feats)
    d = src_feats.shape[-1]
    src_pe = pos_embed_sine(src_xyz, d, scale=pos_scale)
    tgt_pe = pos_embed_sine(tgt_xyz, d, scale=pos_scale)
    src_c, tgt_c = cross_encoder(src_feats, tgt_feats, src_pe, tgt_pe, sd)
    src_corr, tgt_corr, src_ov, tgt_ov = correspondence_decoder(src_c, tgt_c, src_xyz, tgt_xyz, sd, pos_scale)
    L = src_c.shape[0]
    a = torch.cat([src_xyz.expand(L, -1, -1), tgt_corr], dim=1)
    b = torch.cat([src_corr, tgt_xyz.expand(L, -1, -1)], dim=1)
    w = torch.cat([src_ov[..., 0], tgt_ov[..., 0]], dim=1)
    pose = compute_rigid_transform(a, b, w).unsqueeze(1)
    return {
        "src_feats": [src_c], "tgt_feats": [tgt_c],
        "src_kp": [src_xyz], "src_kp_warped": [src_corr],
        "tgt_kp": [tgt_xyz], "tgt_kp_warped": [tgt_corr],
        "src_overlap": [src_ov], "tgt_overlap": [tgt_ov],
        "pose": pose,
    }
