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
                 want_f32=True, want_planes=False, cout=None, ld_out=0, bn_accum=None, tile_list=None,
                 tile_count=None, res_dims=None, out=None):
    """x_planes: (hi, lo) of [g, d, h, w, cin]; w_planes: (hi, lo, scale) of [taps, cout, cin]."""
    x_hi, x_lo = x_planes
    w_hi, w_lo, w_scale = w_planes
    pair = planes == 2
    g, d, h, w, cin = x_hi.shape
    cout = cout or w_hi.shape[1]
    ld = ld_out or cout
    m = g * d * h * w
    dev = x_hi.device
    if out is None:
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
                           bn_accum=bn_accum.data_ptr() if bn_accum is not None else None,
                           tile_list=tile_list.data_ptr() if tile_list is not None else None,
                           tile_count=tile_count.data_ptr() if tile_count is not None else None)
    if res_dims is not None:
        desc.res_d, desc.res_h, desc.res_w = [int(v) for v in res_dims]
    ws = torch.empty(24 << 20, dtype=torch.uint8, device=dev)      # split-K slices (deterministic reduction)
    desc.splitk_ws, desc.splitk_ws_bytes = ws.data_ptr(), ws.numel()
    with _dev(x_hi):
        check(_lib.load().drb_conv3d_igemm(C.byref(desc), stream_ptr()), "drb_conv3d_igemm")
    return out, (o_hi, o_lo)


def error_flag_detail():
    v = (C.c_int * 16)()
    check(_lib.load().drb_error_flag_detail(C.byref(v)), "drb_error_flag_detail")
    return list(v)


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


def batchnorm_small(x, gamma, beta, running_mean, running_var, training, residual=None, relu=False,
                    momentum=0.1, eps=1e-5, want_planes=False, want_stats=False):
    """Same contract as ``batchnorm`` through the one-launch kernel of the deep stages (m <= 1024 rows per grid);
    ``want_stats`` also returns (mean, rstd, scale, shift) [g, c] as drb_bn_save_stats produces them."""
    g, m, c = x.shape
    lib = _lib.load()
    out = torch.empty_like(x)
    o_hi = torch.empty_like(x, dtype=torch.float16) if want_planes else None
    o_lo = torch.empty_like(x, dtype=torch.float16) if want_planes else None
    st = [torch.empty((g, c), dtype=torch.float32, device=x.device) for _ in range(4)] if want_stats else [None] * 4
    with _dev(x):
        check(lib.drb_bn_small(ptr(x), g, m, c, ptr(gamma), ptr(beta), ptr(running_mean), ptr(running_var),
                               int(training), momentum, eps, ptr(residual), int(relu), ptr(out), ptr(o_hi), ptr(o_lo),
                               ptr(st[0]), ptr(st[1]), ptr(st[2]), ptr(st[3]), stream_ptr()), "drb_bn_small")
    return (out, (o_hi, o_lo), tuple(st)) if want_stats else (out, (o_hi, o_lo))


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


# ------------------------------------------------------------------------------------------------
# backward-pass primitives (train_nerf_regtr.py:229)
# ------------------------------------------------------------------------------------------------
def grad_split(x, pair=True, ld_out=None):
    """fp32 [rows, cols] -> (hi, lo, inv_scale device scalar or None): gradient planes with the power-of-two
    pre-scale of the fp16 pair mode."""
    x = x.contiguous().float()
    rows, cols = x.shape
    ld_out = ld_out or cols
    hi = torch.empty((rows, ld_out), dtype=_plane_dtype(pair), device=x.device)
    lo = torch.empty_like(hi) if pair else None
    slot = torch.zeros(2, dtype=torch.float32, device=x.device)
    with _dev(x):
        check(_lib.load().drb_grad_split(ptr(x), rows, cols, cols, ld_out, ptr(hi), ptr(lo), ptr(slot), stream_ptr()),
              "drb_grad_split")
    return hi, lo, (slot[1:] if pair else None)


def conv3d_wgrad(dy_planes, x_planes, k, cout, cin, planes=2, c_real=0, taps_real=0, tile_list=None,
                 tile_count=None, scale=1.0, stage=True):
    """dy_planes: (hi, lo, inv_scale) of [g, d, h, w, cout]; x_planes: (hi, lo) of [g, d, h, w, cin]
    -> dw fp32 [cout, c_real or cin, taps_real or k^3].  stage: hand the kernel a [cout][k^3 cin] workspace
    (vector reductions + one transposition) instead of scattered atomics into dw."""
    dy_hi, dy_lo, inv = dy_planes
    x_hi, x_lo = x_planes
    g, d, h, w, _ = x_hi.shape
    cr, tr = (c_real or cin), (taps_real or k ** 3)
    dw = torch.zeros((cout, cr, tr), dtype=torch.float32, device=x_hi.device)
    ws = torch.empty(cout * k ** 3 * cin, dtype=torch.float32, device=x_hi.device) if stage else None
    desc = _lib.WgradDesc(g=g, d=d, h=h, w=w, cout=cout, cin=cin, kd=k, kh=k, kw=k, planes=planes,
                          dy_hi=dy_hi.data_ptr(), dy_lo=dy_lo.data_ptr() if dy_lo is not None else None,
                          x_hi=x_hi.data_ptr(), x_lo=x_lo.data_ptr() if x_lo is not None else None,
                          scale=scale, scale_dev=inv.data_ptr() if inv is not None else None,
                          dw=dw.data_ptr(), c_real=c_real, taps_real=taps_real,
                          tile_list=tile_list.data_ptr() if tile_list is not None else None,
                          tile_count=tile_count.data_ptr() if tile_count is not None else None,
                          stage=ws.data_ptr() if ws is not None else None,
                          stage_elems=ws.numel() if ws is not None else 0)
    with _dev(x_hi):
        check(_lib.load().drb_conv3d_wgrad(C.byref(desc), stream_ptr()), "drb_conv3d_wgrad")
    return dw


def conv3d_tile_shape(g, d, h, w):
    box, tiles = (C.c_int * 4)(), (C.c_int * 4)()
    check(_lib.load().drb_conv3d_tile_shape(g, d, h, w, C.byref(box), C.byref(tiles)), "drb_conv3d_tile_shape")
    return list(box), list(tiles)


def bn_forward_stats(x, gamma, beta, running_mean, running_var, training, eps=1e-5):
    """-> (mean, rstd, scale, shift) [g, c] exactly as the forward derives them."""
    g, m, c = x.shape
    lib = _lib.load()
    accum = torch.zeros((g, c, 2), dtype=torch.float64, device=x.device)
    outs = [torch.empty((g, c), dtype=torch.float32, device=x.device) for _ in range(4)]
    with _dev(x):
        if training:
            check(lib.drb_bn_stats(ptr(x), g, m, c, ptr(accum), stream_ptr()), "drb_bn_stats")
        check(lib.drb_bn_save_stats(ptr(accum), g, m, c, ptr(gamma), ptr(beta), ptr(running_mean), ptr(running_var),
                                    int(training), eps, ptr(outs[0]), ptr(outs[1]), ptr(outs[2]), ptr(outs[3]),
                                    stream_ptr()), "drb_bn_save_stats")
    return outs


def bn_backward(dy, raw, stats, gamma, training, relu=False, post=None):
    """dy / raw fp32 [g, m, c]; stats from bn_forward_stats -> (dx, dgamma, dbeta); dy is masked in place."""
    g, m, c = raw.shape
    mean, rstd, scale, shift = stats
    sums = torch.empty((g, c, 2), dtype=torch.float64, device=raw.device)
    dx = torch.empty_like(raw)
    dgamma = torch.zeros(c, dtype=torch.float32, device=raw.device)
    dbeta = torch.zeros_like(dgamma)
    with _dev(raw):
        check(_lib.load().drb_bn_backward(ptr(dy), ptr(raw), ptr(post), ptr(scale), ptr(shift), ptr(mean), ptr(rstd),
                                          ptr(gamma), int(relu), int(training), g, m, c, ptr(sums), ptr(dx),
                                          ptr(dgamma), ptr(dbeta), stream_ptr()), "drb_bn_backward")
    return dx, dgamma, dbeta


def maxpool3d_backward(x, dout):
    g, d, h, w, c = x.shape
    dx = torch.empty_like(x)
    with _dev(x):
        check(_lib.load().drb_maxpool3d_backward(ptr(x), ptr(dout), g, d, h, w, c, ptr(dx), stream_ptr()),
              "drb_maxpool3d_backward")
    return dx


def upsample2_add_backward(dsum, coarse_shape):
    g, d, h, w, c = dsum.shape
    _, dc, hc, wc, _ = coarse_shape
    out = torch.empty(coarse_shape, dtype=torch.float32, device=dsum.device)
    with _dev(dsum):
        check(_lib.load().drb_upsample2_add_backward(ptr(dsum), g, d, h, w, c, dc, hc, wc, ptr(out), stream_ptr()),
              "drb_upsample2_add_backward")
    return out


def trilinear_gather_backward(drows, p1_shape, grid_res, mask):
    """drows fp32 [K, c]; -> dp1 [dc, hc, wc, c]."""
    dc, hc, wc, c = p1_shape
    X, Y, Z = grid_res
    dp1 = torch.zeros(p1_shape, dtype=torch.float32, device=drows.device)
    with _dev(drows):
        check(_lib.load().drb_trilinear_gather_backward(ptr(drows), drows.stride(0), 0, dc, hc, wc, c, X, Y, Z,
                                                        ptr(mask), mask.numel(), ptr(dp1), stream_ptr()),
              "drb_trilinear_gather_backward")
    return dp1


def col2im(dcol, x_shape, k, stride, pad, residual=None):
    """dcol fp32 [g, do, ho, wo, kpad] -> dx [g, d, h, w, c]."""
    g, d, h, w, c = x_shape
    kpad = dcol.shape[-1]
    dx = torch.empty(x_shape, dtype=torch.float32, device=dcol.device)
    with _dev(dcol):
        check(_lib.load().drb_col2im(ptr(dcol), g, c, d, h, w, k, stride, pad, kpad, ptr(residual), ptr(dx),
                                     stream_ptr()), "drb_col2im")
    return dx


def layernorm256_backward(x, dy, gamma, dx_accum=None):
    n = x.shape[0]
    dx = dx_accum if dx_accum is not None else torch.empty_like(x)
    dg = torch.zeros(256, dtype=torch.float32, device=x.device)
    db = torch.zeros_like(dg)
    with _dev(x):
        check(_lib.load().drb_layernorm256_backward(ptr(x), ptr(dy), n, ptr(gamma), ptr(dx),
                                                    0 if dx_accum is not None else 1, ptr(dg), ptr(db), stream_ptr()),
              "drb_layernorm256_backward")
    return dx, dg, db


def overlap_sigmoid_backward(feat, ov, dov, w):
    n = feat.shape[0]
    dfeat = torch.zeros_like(feat)
    dw = torch.zeros(256, dtype=torch.float32, device=feat.device)
    db = torch.zeros(1, dtype=torch.float32, device=feat.device)
    with _dev(feat):
        check(_lib.load().drb_overlap_sigmoid_backward(ptr(feat), ptr(ov), ptr(dov), n, ptr(w), ptr(dfeat), ptr(dw),
                                                       ptr(db), stream_ptr()), "drb_overlap_sigmoid_backward")
    return dfeat, dw, db


def softmax_weighted_xyz_backward(s, nk, xyz, dcorr):
    """s fp32 [nq, ld] logits -> dS (new tensor, same shape)."""
    ds = s.clone()
    with _dev(s):
        check(_lib.load().drb_softmax_weighted_xyz_backward(ptr(ds), ds.stride(0), ds.shape[0], nk, ptr(xyz),
                                                            xyz.stride(0), ptr(dcorr), stream_ptr()),
              "drb_softmax_weighted_xyz_backward")
    return ds


def sgemm_strided(A, a_strides, B, b_strides, Cm, c_strides, M, N, K, batch=1, alpha=1.0, accumulate=False):
    with _dev(A):
        check(_lib.load().drb_sgemm_strided(ptr(A), *a_strides, ptr(B), *b_strides, ptr(Cm), *c_strides, M, N, K, batch,
                                            alpha, int(accumulate), stream_ptr()), "drb_sgemm_strided")
    return Cm


def mha_core_backward(q, k, v, dout, heads=8, scale=None):
    """Row matrices [n, 256] (may be strided row views) -> (dq, dk, dv)."""
    nq, nk = q.shape[0], k.shape[0]
    scale = scale if scale is not None else (q.shape[1] // heads) ** -0.5
    lib = _lib.load()
    nbytes = lib.drb_mha_backward_workspace_bytes(nq, nk, heads)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=q.device)
    dq = torch.empty((nq, 256), dtype=torch.float32, device=q.device)
    dk = torch.empty((nk, 256), dtype=torch.float32, device=q.device)
    dv = torch.empty_like(dk)
    with _dev(q):
        check(lib.drb_mha_core_backward(ptr(q), q.stride(0), ptr(k), k.stride(0), ptr(v), v.stride(0), ptr(dout),
                                        dout.stride(0), nq, nk, heads, scale, ptr(dq), 256, ptr(dk), 256, ptr(dv), 256,
                                        ptr(ws), nbytes, stream_ptr()), "drb_mha_core_backward")
    return dq, dk, dv


def procrustes_backward(a, b, w, dpose):
    """a, b [L, n, 3], w [L, n], dpose [L, 3, 4] -> (da, db, dw)."""
    a, b, w, dpose = [t.contiguous().float() for t in (a, b, w, dpose)]
    L, n, _ = a.shape
    da, db, dw = torch.zeros_like(a), torch.zeros_like(b), torch.zeros_like(w)
    with _dev(a):
        check(_lib.load().drb_procrustes_backward(ptr(a), n * 3, ptr(b), n * 3, ptr(w), n, n, None, 0, None, 0, None, 0,
                                                  0, 3, L, ptr(dpose), ptr(da), ptr(db), ptr(dw), None, None, None,
                                                  stream_ptr()), "drb_procrustes_backward")
    return da, db, dw


def hierarchical_downsample_tape(rows, n_src, n_tgt, num_rounds=6, dl0=None, max_total=3000):
    """-> (rows_out, n_src_out, n_tgt_out, tape, [(n_in, n_seg), ...])."""
    lib = _lib.load()
    ld = rows.shape[1]
    n = n_src + n_tgt
    if dl0 is None:
        dl0 = 2.0 * (0.025 * 2.75) / 2.75
    nbytes = lib.drb_downsample_workspace_bytes(n, ld)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=rows.device)
    out = torch.empty_like(rows)
    cap = num_rounds * (2 * n + 2)
    tape = torch.zeros(cap, dtype=torch.int32, device=rows.device)
    info = (C.c_int * (2 * num_rounds))()
    rounds = C.c_int(0)
    a, b = C.c_int(0), C.c_int(0)
    with _dev(rows):
        check(lib.drb_hierarchical_downsample_tape(ptr(rows), n_src, n_tgt, ld, num_rounds, dl0, max_total, ptr(ws),
                                                   nbytes, ptr(out), C.byref(a), C.byref(b), ptr(tape), cap, info,
                                                   C.byref(rounds), stream_ptr()), "drb_hierarchical_downsample_tape")
    return out[:a.value + b.value], a.value, b.value, tape, [(info[2 * r], info[2 * r + 1]) for r in range(rounds.value)]


def downsample_backward(dout, tape, rounds, c):
    """Replays the tape in reverse: dout fp32 [n_out, c] -> gradient of the input rows [n_in0, c]."""
    lib = _lib.load()
    offs, used = [], 0
    for n_in, n_seg in rounds:
        offs.append(used)
        used += n_in + n_seg + 1
    cur = dout.contiguous()
    for (n_in, n_seg), o in reversed(list(zip(rounds, offs))):
        din = torch.empty((n_in, c), dtype=torch.float32, device=dout.device)
        with _dev(dout):
            check(lib.drb_segment_mean_backward(ptr(cur), c, ptr(tape[o:]), ptr(tape[o + n_in:]), n_seg, c, ptr(din), c,
                                                stream_ptr()), "drb_segment_mean_backward")
        cur = din
    return cur


def mha_tc(qkv, split, pairs, heads=8, planes=2, scale=None, want_planes=False):
    """tcgen05 attention over one in_proj output qkv [n, 768] whose rows are two segments [0, split) and
    [split, n); ``pairs`` = list of (q_seg, cross) with q_seg in {0, 1, -1 = both} and cross = attend to the other
    segment; -> fp32 [n, 256] with the rows of every query segment filled (and optionally the 16-bit planes)."""
    n = qkv.shape[0]
    scale = scale if scale is not None else 32 ** -0.5
    lib = _lib.load()
    nbytes = lib.drb_mha_tc_workspace_bytes(n, heads, planes)
    ws = torch.empty(nbytes + 1024, dtype=torch.uint8, device=qkv.device)
    base = (ws.data_ptr() + 1023) // 1024 * 1024
    out = torch.zeros((n, 256), dtype=torch.float32, device=qkv.device)
    o_hi = torch.zeros((n, 256), dtype=_plane_dtype(planes == 2), device=qkv.device) if want_planes else None
    o_lo = torch.zeros((n, 256), dtype=torch.float16, device=qkv.device) if (want_planes and planes == 2) else None
    with _dev(qkv):
        check(lib.drb_mha_tc_pack(ptr(qkv), qkv.stride(0), C.c_void_p(qkv.data_ptr() + 4 * 256), qkv.stride(0),
                                  C.c_void_p(qkv.data_ptr() + 4 * 512), qkv.stride(0), n, split, heads, planes, scale,
                                  C.c_void_p(base), nbytes, stream_ptr()), "drb_mha_tc_pack")
        for q_seg, cross in pairs:
            check(lib.drb_mha_tc_forward(C.c_void_p(base), n, split, heads, planes, q_seg, int(cross), ptr(out), ptr(o_hi),
                                         ptr(o_lo), 256, stream_ptr()), "drb_mha_tc_forward")
    return (out, (o_hi, o_lo)) if want_planes else out
