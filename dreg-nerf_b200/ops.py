"""Thin torch-tensor wrappers over the primitive C-ABI entry points of include/dregb200.h.

Activations are channels-last ``[g, d, h, w, c]`` CUDA tensors; bf16 "planes" are (hi, lo) pairs.
Every function launches on the current CUDA stream of the tensor's device.  No fallbacks.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import check, ptr, stream_ptr


def _dev(t):
    if not t.is_cuda:
        raise _lib.DrbError("libdregb200 ops need CUDA tensors")
    return torch.cuda.device(t.device)


def _plane_dtype(pair):
    return torch.float16 if pair else torch.bfloat16


def split_planes(x, want_lo=True):
    """fp32 -> (hi, lo): an fp16 pair when want_lo, else one bf16 plane (lo None)."""
    x = x.contiguous().float()
    hi = torch.empty_like(x, dtype=_plane_dtype(want_lo))
    lo = torch.empty_like(x, dtype=torch.float16) if want_lo else None
    with _dev(x):
        check(_lib.load().drb_split_planes(ptr(x), ptr(hi), ptr(lo), x.numel(), stream_ptr()), "drb_split_planes")
    return hi, lo


def weight_scale(w):
    """Power-of-two pre-scale for the fp16 pair mode (max|w| -> [256, 512))."""
    s = C.c_float(1.0)
    with _dev(w):
        check(_lib.load().drb_weight_scale(ptr(w), w.numel(), C.byref(s), stream_ptr()), "drb_weight_scale")
    return s.value


def pack_conv_weight(w, cin_pad=None, pair=True):
    """torch Conv3d / Linear weight [cout, cin, *k] -> (hi, lo, scale), planes [taps, cout, cin_pad]."""
    w = w.contiguous().float()
    cout, cin = w.shape[0], w.shape[1]
    taps = w.numel() // (cout * cin)
    cin_pad = cin_pad or cin
    scale = weight_scale(w) if pair else 1.0
    hi = torch.empty((taps, cout, cin_pad), dtype=_plane_dtype(pair), device=w.device)
    lo = torch.empty_like(hi) if pair else None
    with _dev(w):
        check(_lib.load().drb_pack_conv_weight(ptr(w), cout, cin, taps, cin_pad, scale, ptr(hi), ptr(lo),
                                               stream_ptr()), "drb_pack_conv_weight")
    return hi, lo, scale


def pack_conv_weight_im2col(w, kpad, pair=True):
    w = w.contiguous().float()
    cout, cin = w.shape[0], w.shape[1]
    taps = w.numel() // (cout * cin)
    scale = weight_scale(w) if pair else 1.0
    hi = torch.empty((1, cout, kpad), dtype=_plane_dtype(pair), device=w.device)
    lo = torch.empty_like(hi) if pair else None
    with _dev(w):
        check(_lib.load().drb_pack_conv_weight_im2col(ptr(w), cout, cin, taps, kpad, scale, ptr(hi), ptr(lo),
                                                      stream_ptr()), "drb_pack_conv_weight_im2col")
    return hi, lo, scale


def conv3d_igemm(x_planes, w_planes, k, planes=2, bias=None, residual=None, relu=False, out_scale=1.0,
                 want_f32=True, want_planes=False, cout=None, ld_out=0, bn_accum=None):
    """x_planes: (hi, lo) of [g, d, h, w, cin]; w_planes: (hi, lo, scale) of [taps, cout, cin]."""
    x_hi, x_lo = x_planes
    w_hi, w_lo, w_scale = w_planes
    pair = planes == 2
    g, d, h, w, cin = x_hi.shape
    cout = cout or w_hi.shape[1]
    ld = ld_out or cout
    m = g * d * h * w
    dev = x_hi.device
    out = torch.empty((m, ld), dtype=torch.float32, device=dev) if want_f32 else None
    o_hi = torch.empty((m, ld), dtype=_plane_dtype(pair), device=dev) if want_planes else None
    o_lo = torch.empty((m, ld), dtype=torch.float16, device=dev) if (want_planes and pair) else None
    desc = _lib.Conv3dDesc(g=g, d=d, h=h, w=w, cin=cin, cout=cout, kd=k, kh=k, kw=k, planes=planes,
                           relu=int(relu), acc_scale=1.0 / w_scale, out_scale=out_scale,
                           x_hi=x_hi.data_ptr(), x_lo=x_lo.data_ptr() if x_lo is not None else None,
                           w_hi=w_hi.data_ptr(), w_lo=w_lo.data_ptr() if w_lo is not None else None,
                           bias=bias.data_ptr() if bias is not None else None,
                           residual=residual.data_ptr() if residual is not None else None,
                           out=out.data_ptr() if out is not None else None,
                           out_hi=o_hi.data_ptr() if o_hi is not None else None,
                           out_lo=o_lo.data_ptr() if o_lo is not None else None, ld_out=ld,
                           bn_accum=bn_accum.data_ptr() if bn_accum is not None else None)
    with _dev(x_hi):
        check(_lib.load().drb_conv3d_igemm(C.byref(desc), stream_ptr()), "drb_conv3d_igemm")
    return out, (o_hi, o_lo)


def igemm_error_flag():
    v = C.c_int(0)
    check(_lib.load().drb_igemm_error_flag(C.byref(v)), "drb_igemm_error_flag")
    return v.value


def im2col(x, k, stride, pad, kpad, channel_slice=None):
    """x: fp32 [g, c, d, h, w] view with arbitrary strides (torch NCDHW indexing) -> planes
    [g, do, ho, wo, kpad]."""
    g, c, d, h, w = x.shape
    od, oh, ow = [(n + 2 * pad - k) // stride + 1 for n in (d, h, w)]
    hi = torch.empty((g, od, oh, ow, kpad), dtype=torch.float16, device=x.device)
    lo = torch.empty_like(hi)
    desc = _lib.Im2colDesc(x=x.data_ptr(), sg=x.stride(0), sc=x.stride(1), sd=x.stride(2), sh=x.stride(3),
                           sw=x.stride(4), g=g, c=c, d=d, h=h, w=w, k=k, stride=stride, pad=pad, kpad=kpad)
    with _dev(x):
        check(_lib.load().drb_im2col(C.byref(desc), ptr(hi), ptr(lo), stream_ptr()), "drb_im2col")
    return hi, lo


def im2col_stem(x):
    """x: fp32 [1, 4, d, h, w] strided view -> planes [1, do, ho, wo, 512] (conv1 fast path)."""
    _, c, d, h, w = x.shape
    assert c == 4
    od, oh, ow = [(n + 4 - 5) // 2 + 1 for n in (d, h, w)]
    hi = torch.empty((1, od, oh, ow, 512), dtype=torch.float16, device=x.device)
    lo = torch.empty_like(hi)
    scratch = torch.empty(d * h * w * 4, dtype=torch.float32, device=x.device)
    with _dev(x):
        check(_lib.load().drb_im2col_stem(ptr(x), x.stride(1), x.stride(2), x.stride(3), x.stride(4), d, h, w,
                                          ptr(scratch), ptr(hi), ptr(lo), stream_ptr()), "drb_im2col_stem")
    return hi, lo


def batchnorm(x, gamma, beta, running_mean, running_var, training, residual=None, relu=False,
              momentum=0.1, eps=1e-5, want_planes=False):
    """x fp32 [g, m, c] -> y (and planes).  Updates running buffers in place when training."""
    g, m, c = x.shape
    lib = _lib.load()
    accum = torch.empty((g, c, 2), dtype=torch.float64, device=x.device)
    scale = torch.empty((g, c), dtype=torch.float32, device=x.device)
    shift = torch.empty_like(scale)
    out = torch.empty_like(x)
    o_hi = torch.empty_like(x, dtype=torch.float16) if want_planes else None
    o_lo = torch.empty_like(x, dtype=torch.float16) if want_planes else None
    with _dev(x):
        if training:
            check(lib.drb_bn_stats(ptr(x), g, m, c, ptr(accum), stream_ptr()), "drb_bn_stats")
        check(lib.drb_bn_finalize(ptr(accum), g, m, c, ptr(gamma), ptr(beta), ptr(running_mean),
                                  ptr(running_var), int(training), momentum, eps, ptr(scale), ptr(shift),
                                  stream_ptr()), "drb_bn_finalize")
        check(lib.drb_scale_shift_act(ptr(x), ptr(scale), ptr(shift), ptr(residual), int(relu), g, m, c,
                                      ptr(out), ptr(o_hi), ptr(o_lo), stream_ptr()), "drb_scale_shift_act")
    return out, (o_hi, o_lo)


def batchnorm_fused(x, gamma, beta, running_mean, running_var, training, residual=None, relu=False,
                    momentum=0.1, eps=1e-5, want_planes=False):
    """Same contract as ``batchnorm`` through the single-launch drb_bn_apply."""
    g, m, c = x.shape
    lib = _lib.load()
    accum = torch.empty((g, c, 2), dtype=torch.float64, device=x.device)
    out = torch.empty_like(x)
    o_hi = torch.empty_like(x, dtype=torch.float16) if want_planes else None
    o_lo = torch.empty_like(x, dtype=torch.float16) if want_planes else None
    with _dev(x):
        if training:
            check(lib.drb_bn_stats(ptr(x), g, m, c, ptr(accum), stream_ptr()), "drb_bn_stats")
        check(lib.drb_bn_apply(ptr(x), ptr(accum), g, m, c, ptr(gamma), ptr(beta), ptr(running_mean),
                               ptr(running_var), int(training), momentum, eps, ptr(residual), int(relu), ptr(out),
                               ptr(o_hi), ptr(o_lo), stream_ptr()), "drb_bn_apply")
    return out, (o_hi, o_lo)


def maxpool3d(x):
    g, d, h, w, c = x.shape
    od, oh, ow = (d - 1) // 2 + 1, (h - 1) // 2 + 1, (w - 1) // 2 + 1
    out = torch.empty((g, od, oh, ow, c), dtype=torch.float32, device=x.device)
    with _dev(x):
        check(_lib.load().drb_maxpool3d(ptr(x), g, d, h, w, c, ptr(out), None, None, stream_ptr()), "drb_maxpool3d")
    return out


def upsample2_add(coarse, lateral):
    g, d, h, w, c = lateral.shape
    _, dc, hc, wc, _ = coarse.shape
    out = torch.empty_like(lateral)
    with _dev(lateral):
        check(_lib.load().drb_upsample2_add(ptr(coarse), dc, hc, wc, ptr(lateral), g, d, h, w, c, ptr(out), None,
                                            None, stream_ptr()), "drb_upsample2_add")
    return out


def trilinear_gather(p1, grid, mask):
    """p1 fp32 [dc, hc, wc, c]; grid the reference's [1, 7, Z, X, Y] view; mask int64 [K]
    -> rows [K, 4 + c] = [x y z 0 | features]."""
    dc, hc, wc, c = p1.shape
    _, _, Z, X, Y = grid.shape
    k = mask.numel()
    rows = torch.empty((k, 4 + c), dtype=torch.float32, device=p1.device)
    with _dev(p1):
        check(_lib.load().drb_trilinear_gather(ptr(p1), dc, hc, wc, c, ptr(grid), grid.stride(1), grid.stride(2),
                                               grid.stride(3), grid.stride(4), X, Y, Z, ptr(mask), k, ptr(rows),
                                               4 + c, stream_ptr()), "drb_trilinear_gather")
    return rows


def hierarchical_downsample(rows, n_src, n_tgt, num_rounds=6, dl0=None, max_total=3000):
    lib = _lib.load()
    ld = rows.shape[1]
    n = n_src + n_tgt
    if dl0 is None:
        dl0 = 2.0 * (0.025 * 2.75) / 2.75
    nbytes = lib.drb_downsample_workspace_bytes(n, ld)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=rows.device)
    out = torch.empty_like(rows)
    a, b = C.c_int(0), C.c_int(0)
    with _dev(rows):
        check(lib.drb_hierarchical_downsample(ptr(rows), n_src, n_tgt, ld, num_rounds, dl0, max_total, ptr(ws),
                                              nbytes, ptr(out), C.byref(a), C.byref(b), stream_ptr()),
              "drb_hierarchical_downsample")
    return out[:a.value + b.value], a.value, b.value


def pos_embed_sine(xyz, scale=1.0):
    xyz = xyz.contiguous().float()
    n = xyz.shape[0]
    out = torch.empty((n, 256), dtype=torch.float32, device=xyz.device)
    with _dev(xyz):
        check(_lib.load().drb_pos_embed_sine(ptr(xyz), xyz.shape[1], n, scale, ptr(out), stream_ptr()),
              "drb_pos_embed_sine")
    return out


def layernorm256(x, gamma, beta, add=None, want_planes=False):
    n = x.shape[0]
    out = torch.empty_like(x)
    o_hi = torch.empty_like(x, dtype=torch.float16) if want_planes else None
    o_lo = torch.empty_like(x, dtype=torch.float16) if want_planes else None
    with _dev(x):
        check(_lib.load().drb_layernorm256(ptr(x), n, ptr(gamma), ptr(beta), ptr(add), ptr(out), ptr(o_hi),
                                           ptr(o_lo), stream_ptr()), "drb_layernorm256")
    return out, (o_hi, o_lo)


def mha_core(q, k, v, heads=8, scale=None):
    """q [nq, 256], k / v [nk, 256] (may be strided row views) -> [nq, 256]."""
    nq, nk = q.shape[0], k.shape[0]
    scale = scale if scale is not None else (q.shape[1] // heads) ** -0.5
    out = torch.empty((nq, q.shape[1]), dtype=torch.float32, device=q.device)
    with _dev(q):
        check(_lib.load().drb_mha_core(ptr(q), q.stride(0), ptr(k), k.stride(0), ptr(v), v.stride(0), nq, nk,
                                       heads, scale, ptr(out), None, None, out.stride(0), stream_ptr()),
              "drb_mha_core")
    return out


def softmax_weighted_xyz(s, nk, xyz):
    nq = s.shape[0]
    out = torch.empty((nq, 3), dtype=torch.float32, device=s.device)
    with _dev(s):
        check(_lib.load().drb_softmax_weighted_xyz(ptr(s), s.stride(0), nq, nk, ptr(xyz), xyz.stride(0), ptr(out),
                                                   stream_ptr()), "drb_softmax_weighted_xyz")
    return out


def overlap_sigmoid(feat, w, b):
    n = feat.shape[0]
    out = torch.empty(n, dtype=torch.float32, device=feat.device)
    with _dev(feat):
        check(_lib.load().drb_overlap_sigmoid(ptr(feat), n, ptr(w), ptr(b), ptr(out), stream_ptr()),
              "drb_overlap_sigmoid")
    return out


def procrustes(a, b, w):
    """a, b [L, n, 3], w [L, n] -> [L, 3, 4] (se3.py:89-140)."""
    a, b, w = a.contiguous().float(), b.contiguous().float(), w.contiguous().float()
    L, n, _ = a.shape
    out = torch.empty((L, 3, 4), dtype=torch.float32, device=a.device)
    with _dev(a):
        check(_lib.load().drb_procrustes(ptr(a), n * 3, ptr(b), n * 3, ptr(w), n, n, None, 0, None, 0, None, 0, 0,
                                         3, L, ptr(out), stream_ptr()), "drb_procrustes")
    return out
