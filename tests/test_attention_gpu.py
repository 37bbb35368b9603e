"""The tcgen05 FlashAttention-style kernel (csrc/attention.cu) against fp64 torch attention, and the whole forward
with it switched on against the oracle."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref_attention(q, k, v, heads=8):
    nq, nk = q.shape[0], k.shape[0]
    qh = q.double().view(nq, heads, 32).transpose(0, 1) / math.sqrt(32)
    kh = k.double().view(nk, heads, 32).transpose(0, 1)
    vh = v.double().view(nk, heads, 32).transpose(0, 1)
    att = torch.softmax(qh @ kh.transpose(1, 2), dim=-1)
    return (att @ vh).transpose(0, 1).reshape(nq, heads * 32)


@pytest.mark.parametrize("planes,tol", [(2, 2e-5), (1, 2e-2)])
@pytest.mark.parametrize("ns,nt", [(333, 257), (128, 128), (1100, 900), (5, 700)])
def test_mha_tc_matches_torch(pkg, cuda, planes, tol, ns, nt):
    from importlib import import_module
    ops = import_module("dreg-nerf_b200.ops")
    torch.manual_seed(ns + nt)
    m = ns + nt
    qkv = torch.randn(m, 768) * 1.5
    q, k, v = qkv[:, :256], qkv[:, 256:512], qkv[:, 512:]
    # self-attention of both clouds, then both cross directions (transformer.py:240-281)
    want_self = torch.cat([_ref_attention(q[:ns], k[:ns], v[:ns]), _ref_attention(q[ns:], k[ns:], v[ns:])])
    want_cross = torch.cat([_ref_attention(q[:ns], k[ns:], v[ns:]), _ref_attention(q[ns:], k[:ns], v[:ns])])
    dq = qkv.to(cuda)
    got_self, (hi, lo) = ops.mha_tc(dq, ns, [(-1, 0)], planes=planes, want_planes=True)     # both clouds, one launch
    got_cross = ops.mha_tc(dq, ns, [(0, 1), (1, 1)], planes=planes)                          # one direction per launch
    flag = ops.igemm_error_flag()          # synchronises; readable even after a device-side watchdog trap
    assert flag == 0, "pipeline watchdog code %d, sites %s" % (flag, ops.error_flag_detail())
    e1 = ((got_self.cpu().double() - want_self).abs().max() / want_self.abs().max()).item()
    e2 = ((got_cross.cpu().double() - want_cross).abs().max() / want_cross.abs().max()).item()
    print("tcgen05 attention ns=%d nt=%d planes=%d: self %.2e cross %.2e" % (ns, nt, planes, e1, e2))
    assert e1 < tol and e2 < tol
    rec = hi.float() + (lo.float() if lo is not None else 0)
    assert ((rec.cpu().double() - want_self).abs().max() / want_self.abs().max()).item() < max(tol, 1e-2 if planes == 1 else tol)


def test_mha_tc_long_sequence(pkg, cuda):
    """BASELINE.json configs[4] size: ~8k tokens per cloud."""
    from importlib import import_module
    ops = import_module("dreg-nerf_b200.ops")
    torch.manual_seed(1)
    n = 8192
    qkv = torch.randn(2 * n, 768)
    want = _ref_attention(qkv[:n, :256], qkv[n:, 256:512], qkv[n:, 512:])
    got = ops.mha_tc(qkv.to(cuda), n, [(0, 1)], planes=2)[:n]           # source queries against the target cloud
    err = ((got.cpu().double() - want).abs().max() / want.abs().max()).item()
    print("tcgen05 attention 8192 x 8192: rel err %.2e" % err)
    assert err < 2e-5


def test_forward_with_tc_attention(pkg, cuda):
    from oracle import regtr
    torch.manual_seed(0)
    model = pkg.NeRFRegTr()
    sd = pkg.synthetic.seeded_state_dict(model, seed=0, attn_gain=4.0)
    model.load_state_dict(sd)
    model = model.to(cuda).eval()
    model.tc_attention = True
    data = pkg.synthetic.make_pair(res=32, pair_id=0)
    with torch.no_grad():
        out = model(pkg.synthetic.to_device(data, cuda))
        ref = regtr.forward(sd, data, training=False)
    rel = lambda a, b: ((a.cpu().double() - b.double()).abs().max() / b.double().abs().max()).item()
    errs = {k: rel(out[k][0], ref[k][0]) for k in ("src_feats", "tgt_feats", "src_kp_warped", "tgt_kp_warped", "src_overlap")}
    errs["pose"] = rel(out["pose"], ref["pose"])
    print("forward with tcgen05 attention:", {k: "%.2e" % v for k, v in errs.items()})
    assert max(errs.values()) < 1e-3
