"""Per-kernel parity against torch CPU references (fp64 where cheap), through the C ABI."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _ops():
    import importlib
    return importlib.import_module("dreg-nerf_b200.ops")


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()


def test_split_planes(cuda):
    ops = _ops()
    g = torch.Generator().manual_seed(0)
    x = torch.randn(100003, generator=g) * torch.logspace(-3, 3, 100003)
    x = x.clamp(-6e4, 6e4)
    hi, lo = ops.split_planes(x.to(cuda))
    assert hi.dtype == torch.float16 and lo.dtype == torch.float16
    rec = hi.float() + lo.float()
    # 22 significand bits inside fp16's normal range, 2^-24 absolute below it
    assert ((rec.cpu() - x).abs() <= x.abs() * 2.0 ** -21 + 2.0 ** -24).all()
    hi1, lo1 = ops.split_planes(x.to(cuda), want_lo=False)
    assert lo1 is None and hi1.dtype == torch.bfloat16
    assert ((hi1.float().cpu() - x).abs() <= x.abs() * 2.0 ** -8).all()


@pytest.mark.parametrize("planes,tol", [(2, 2e-6), (1, 2e-2)])
@pytest.mark.parametrize("m,cin,cout", [(300, 256, 768), (128, 64, 64), (1000, 1024, 256), (77, 256, 200)])
def test_igemm_linear(cuda, planes, tol, m, cin, cout):
    """nn.Linear as the 1x1x1 case, ragged M and ragged N."""
    ops = _ops()
    g = torch.Generator().manual_seed(m + cin + cout)
    x = torch.randn(m, cin, generator=g)
    w = torch.randn(cout, cin, generator=g) / math.sqrt(cin)
    b = torch.randn(cout, generator=g)
    ld = (cout + 7) // 8 * 8
    xp = ops.split_planes(x.to(cuda).view(1, 1, 1, m, cin), want_lo=planes == 2)
    wp = ops.pack_conv_weight(w.to(cuda), pair=planes == 2)
    out, _ = ops.conv3d_igemm(xp, wp, 1, planes=planes, bias=b.to(cuda), ld_out=ld)
    torch.cuda.synchronize()
    assert ops.igemm_error_flag() == 0
    ref = x.double() @ w.double().t() + b.double()
    assert _rel(out[:, :cout], ref) < tol


@pytest.mark.parametrize("shape", [(2, 8, 8, 8), (1, 6, 10, 12), (2, 4, 4, 4), (1, 2, 2, 2), (1, 16, 16, 16)])
@pytest.mark.parametrize("cin,cout,k", [(64, 64, 3), (128, 256, 3), (64, 256, 1)])
def test_igemm_conv3d(cuda, shape, cin, cout, k):
    """nn.Conv3d stride 1 'same' through 5-D TMA boxes with zero-filled halos."""
    ops = _ops()
    g, d, h, w = shape
    gen = torch.Generator().manual_seed(d * 100 + cin + k)
    x = torch.randn(g, cin, d, h, w, generator=gen)
    wt = torch.randn(cout, cin, k, k, k, generator=gen) / math.sqrt(cin * k ** 3)
    bias = torch.randn(cout, generator=gen)
    xp = ops.split_planes(x.permute(0, 2, 3, 4, 1).contiguous().to(cuda))
    wp = ops.pack_conv_weight(wt.to(cuda))
    out, _ = ops.conv3d_igemm(xp, wp, k, planes=2, bias=bias.to(cuda))
    torch.cuda.synchronize()
    assert ops.igemm_error_flag() == 0
    ref = F.conv3d(x.double(), wt.double(), bias.double(), padding=k // 2).permute(0, 2, 3, 4, 1).reshape(-1, cout)
    assert _rel(out, ref) < 2e-6


def test_igemm_two_tile_bf16(cuda):
    """The bf16 kernel with two M tiles per work item (igemm2_kernel: large volumes, 256-wide weight tile): plain
    epilogue, bias + ReLU + planes, the FPN's coarse residual, and an odd-length tile list - against torch on the
    same bf16-rounded operands."""
    ops = _ops()
    g, d, h, w, cin, cout, k = 2, 34, 35, 33, 64, 256, 3
    gen = torch.Generator().manual_seed(5)
    x = torch.randn(g, cin, d, h, w, generator=gen)
    wt = torch.randn(cout, cin, k, k, k, generator=gen) / math.sqrt(cin * k ** 3)
    bias = torch.randn(cout, generator=gen)
    top = torch.randn(g, cout, (d + 1) // 2, (h + 1) // 2, (w + 1) // 2, generator=gen)
    r16 = lambda t: t.to(torch.bfloat16).float()
    ref = F.conv3d(r16(x), r16(wt), bias, padding=1)                       # fp32 accumulate of exact bf16 products
    up = F.interpolate(top, scale_factor=2)[:, :, :d, :h, :w]
    cl = lambda t: t.permute(0, 2, 3, 4, 1).reshape(-1, cout)
    xp = ops.split_planes(x.permute(0, 2, 3, 4, 1).contiguous().to(cuda), want_lo=False)
    wp = ops.pack_conv_weight(wt.to(cuda), pair=False)
    box, tiles = ops.conv3d_tile_shape(g, d, h, w)
    n_tiles = tiles[0] * tiles[1] * tiles[2] * tiles[3]
    assert n_tiles >= 4 * 148, "the shape must be large enough to select the two-tile kernel"
    out, _ = ops.conv3d_igemm(xp, wp, k, planes=1, bias=bias.to(cuda))
    assert _rel(out, cl(ref)) < 2e-5
    out, (hi, _) = ops.conv3d_igemm(xp, wp, k, planes=1, bias=bias.to(cuda), relu=True, want_planes=True)
    assert _rel(out, cl(torch.relu(ref))) < 2e-5 and _rel(hi.float(), cl(torch.relu(ref))) < 5e-3
    res_cl = top.permute(0, 2, 3, 4, 1).contiguous().to(cuda)
    out, _ = ops.conv3d_igemm(xp, wp, k, planes=1, bias=bias.to(cuda), residual=res_cl, res_dims=top.shape[2:])
    assert _rel(out, cl(ref + up)) < 2e-5
    # output-sparse: an odd number of listed tiles, the others keep what they held
    pick = torch.randperm(n_tiles, generator=gen)[: n_tiles // 2 * 2 - 1].sort().values.int()
    keep = torch.full((g * d * h * w, cout), 7.0, device=cuda)
    out, _ = ops.conv3d_igemm(xp, wp, k, planes=1, bias=bias.to(cuda), tile_list=pick.to(cuda),
                              tile_count=torch.tensor([pick.numel()], dtype=torch.int32, device=cuda), out=keep)
    tile_of = torch.zeros(g, d, h, w, dtype=torch.long)
    ig, iz, iy, ix = torch.meshgrid(torch.arange(g), torch.arange(d), torch.arange(h), torch.arange(w), indexing="ij")
    tile_of = ((ig // box[0] * tiles[1] + iz // box[1]) * tiles[2] + iy // box[2]) * tiles[3] + ix // box[3]
    listed = torch.isin(tile_of.reshape(-1), pick.long())
    torch.cuda.synchronize()
    assert ops.igemm_error_flag() == 0
    assert _rel(out.cpu()[listed], cl(ref)[listed]) < 2e-5
    assert bool((out.cpu()[~listed] == 7.0).all())


@pytest.mark.parametrize("shape,cin,cout,k", [((2, 4, 4, 4), 512, 512, 3), ((2, 8, 8, 8), 256, 256, 3),
                                              ((1, 1, 1, 90), 2048, 256, 1)])
def test_igemm_split_k(cuda, shape, cin, cout, k):
    """Deep backbone layers: 1-8 output tiles, K up to 13824 -> split-K work items + atomic reduction."""
    ops = _ops()
    g, d, h, w = shape
    gen = torch.Generator().manual_seed(cin + cout)
    x = torch.randn(g, cin, d, h, w, generator=gen)
    wt = torch.randn(cout, cin, k, k, k, generator=gen) / math.sqrt(cin * k ** 3)
    bias = torch.randn(cout, generator=gen)
    xp = ops.split_planes(x.permute(0, 2, 3, 4, 1).contiguous().to(cuda))
    wp = ops.pack_conv_weight(wt.to(cuda))
    out, _ = ops.conv3d_igemm(xp, wp, k, planes=2, bias=bias.to(cuda))
    torch.cuda.synchronize()
    assert ops.igemm_error_flag() == 0
    ref = F.conv3d(x.double(), wt.double(), bias.double(), padding=k // 2).permute(0, 2, 3, 4, 1).reshape(-1, cout)
    assert _rel(out, ref) < 2e-6


@pytest.mark.parametrize("shape,cin,cout", [((2, 8, 8, 8), 64, 256), ((2, 6, 10, 12), 128, 64), ((2, 4, 4, 4), 512, 128)])
def test_igemm_fused_bn_sums(cuda, shape, cin, cout):
    """Per-channel sum / sum of squares of the conv output from the GEMM epilogue (or the fallback pass
    when the reduction is split or a tile straddles two grids)."""
    ops = _ops()
    g, d, h, w = shape
    gen = torch.Generator().manual_seed(cin * 3 + cout)
    x = torch.randn(g, cin, d, h, w, generator=gen)
    wt = torch.randn(cout, cin, 3, 3, 3, generator=gen) / math.sqrt(cin * 27)
    xp = ops.split_planes(x.permute(0, 2, 3, 4, 1).contiguous().to(cuda))
    wp = ops.pack_conv_weight(wt.to(cuda))
    accum = torch.full((g, cout, 2), 7.0, dtype=torch.float64, device=cuda)
    out, _ = ops.conv3d_igemm(xp, wp, 3, planes=2, bn_accum=accum)
    torch.cuda.synchronize()
    ref = F.conv3d(x.double(), wt.double(), padding=1)
    assert _rel(out, ref.permute(0, 2, 3, 4, 1).reshape(-1, cout)) < 2e-6
    assert _rel(accum[:, :, 0], ref.sum(dim=(2, 3, 4))) < 1e-5
    assert _rel(accum[:, :, 1], (ref * ref).sum(dim=(2, 3, 4))) < 1e-5


def test_igemm_epilogue(cuda):
    """out = relu((acc + bias) * scale + residual), fp32 and plane outputs."""
    ops = _ops()
    gen = torch.Generator().manual_seed(5)
    m, cin, cout = 500, 256, 256
    x = torch.randn(m, cin, generator=gen)
    w = torch.randn(cout, cin, generator=gen) / 16
    b = torch.randn(cout, generator=gen)
    res = torch.randn(m, cout, generator=gen)
    xp = ops.split_planes(x.to(cuda).view(1, 1, 1, m, cin))
    wp = ops.pack_conv_weight(w.to(cuda))
    out, (ohi, olo) = ops.conv3d_igemm(xp, wp, 1, bias=b.to(cuda), residual=res.to(cuda), relu=True,
                                       out_scale=0.25, want_planes=True)
    torch.cuda.synchronize()
    ref = torch.relu((x.double() @ w.double().t() + b.double()) * 0.25 + res.double())
    assert _rel(out, ref) < 2e-6
    assert _rel(ohi.float() + olo.float(), ref) < 4e-6


@pytest.mark.parametrize("cin,cout,k,stride,size,extra", [(4, 64, 5, 2, 16, 3), (128, 128, 3, 2, 8, 3), (256, 512, 1, 2, 8, 3),
                                                            (128, 128, 3, 2, 9, 0), (256, 512, 1, 2, 8, 0), (64, 64, 3, 2, 12, 0)])
def test_im2col_conv(cuda, cin, cout, k, stride, size, extra):
    """conv1 (5^3 s2, Cin 4) and the stride-2 convs via im2col + 1x1x1 igemm; strided input view (extra = 3: the
    element-per-thread kernel) and plain channels-last storage (extra = 0: the 8-channels-per-thread fast path the
    engine's strided convolutions take)."""
    ops = _ops()
    gen = torch.Generator().manual_seed(cin + k)
    base = torch.randn(2, size, size, size, cin + extra, generator=gen)   # [g, X, Y, Z, C] storage
    x = base.permute(0, 4, 3, 1, 2)[:, extra:]                             # [g, C, Z, X, Y] strided view
    wt = torch.randn(cout, cin, k, k, k, generator=gen) / math.sqrt(cin * k ** 3)
    kpad = (cin * k ** 3 + 63) // 64 * 64
    xd = base.to(cuda).permute(0, 4, 3, 1, 2)[:, extra:]
    cols = ops.im2col(xd, k, stride, k // 2, kpad)
    wp = ops.pack_conv_weight_im2col(wt.to(cuda), kpad)
    out, _ = ops.conv3d_igemm(cols, wp, 1)
    torch.cuda.synchronize()
    ref = F.conv3d(x.double(), wt.double(), stride=stride, padding=k // 2).permute(0, 2, 3, 4, 1).reshape(-1, cout)
    assert _rel(out, ref) < 2e-6


def test_im2col_stem_matches_generic(cuda):
    """The stem fast path (pack + warp-per-voxel im2col) writes exactly what the generic kernel does."""
    ops = _ops()
    gen = torch.Generator().manual_seed(41)
    base = torch.randn(1, 20, 18, 22, 7, generator=gen)                    # [g, X, Y, Z, C] storage
    xd = base.to(cuda).permute(0, 4, 3, 1, 2)[:, 3:]                        # [1, 4, Z, X, Y] strided view
    a_hi, a_lo = ops.im2col(xd, 5, 2, 2, 512)
    b_hi, b_lo = ops.im2col_stem(xd)
    assert torch.equal(a_hi, b_hi) and torch.equal(a_lo, b_lo)


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("training", [True, False])
def test_batchnorm(cuda, training, fused):
    ops = _ops()
    gen = torch.Generator().manual_seed(3)
    g, m, c = 2, 1000, 128
    x = torch.randn(g, m, c, generator=gen) * 3 + 1.5
    gamma, beta = torch.rand(c, generator=gen) + 0.5, torch.randn(c, generator=gen)
    rm, rv = torch.randn(c, generator=gen) * 0.1, torch.rand(c, generator=gen) + 0.5
    res = torch.randn(g, m, c, generator=gen)
    rm_d, rv_d = rm.clone().to(cuda), rv.clone().to(cuda)
    fn = ops.batchnorm_fused if fused else ops.batchnorm
    out, (hi, lo) = fn(x.to(cuda), gamma.to(cuda), beta.to(cuda), rm_d, rv_d, training,
                       residual=res.to(cuda), relu=True, want_planes=True)
    torch.cuda.synchronize()
    rm_ref, rv_ref = rm.clone(), rv.clone()
    refs = []
    for gi in range(g):   # the reference runs the grids as successive B = 1 calls
        xi = x[gi].t().reshape(1, c, m, 1, 1)
        yi = F.batch_norm(xi, rm_ref, rv_ref, gamma, beta, training, 0.1, 1e-5)
        refs.append(torch.relu(yi.reshape(c, m).t() + res[gi]))
    ref = torch.stack(refs)
    assert _rel(out, ref) < 1e-5
    assert _rel(hi.float() + lo.float(), ref) < 4e-6
    assert _rel(rm_d, rm_ref) < 1e-5 and _rel(rv_d, rv_ref) < 1e-5


@pytest.mark.parametrize("training", [True, False])
@pytest.mark.parametrize("g,m,c", [(2, 512, 256), (2, 64, 2048), (2, 8, 512), (3, 1000, 36)])
def test_batchnorm_small_matches_general_path(cuda, training, g, m, c):
    """The one-launch BatchNorm of the deep stages against torch and against the general three-launch path,
    including the statistics the backward pass keeps."""
    ops = _ops()
    gen = torch.Generator().manual_seed(g * m + c)
    x = torch.randn(g, m, c, generator=gen) * 3 + 1.5
    gamma, beta = torch.rand(c, generator=gen) + 0.5, torch.randn(c, generator=gen)
    rm, rv = torch.randn(c, generator=gen) * 0.1, torch.rand(c, generator=gen) + 0.5
    res = torch.randn(g, m, c, generator=gen)
    dev = lambda t: t.clone().to(cuda)
    rm_a, rv_a, rm_b, rv_b = dev(rm), dev(rv), dev(rm), dev(rv)
    out, (hi, lo), st = ops.batchnorm_small(dev(x), dev(gamma), dev(beta), rm_a, rv_a, training, residual=dev(res),
                                            relu=True, want_planes=True, want_stats=True)
    ref_out, (rhi, rlo) = ops.batchnorm_fused(dev(x), dev(gamma), dev(beta), rm_b, rv_b, training, residual=dev(res),
                                              relu=True, want_planes=True)
    want = ops.bn_forward_stats(dev(x), dev(gamma), dev(beta), dev(rm), dev(rv), training)
    torch.cuda.synchronize()
    assert _rel(out, ref_out) < 2e-6 and _rel(rm_a, rm_b) < 1e-6 and _rel(rv_a, rv_b) < 1e-6
    assert _rel(hi.float() + lo.float(), rhi.float() + rlo.float()) < 4e-6
    for a, b in zip(st, want):
        assert _rel(a, b) < 2e-6
    rm_ref, rv_ref = rm.clone(), rv.clone()
    refs = []
    for gi in range(g):
        xi = x[gi].t().reshape(1, c, m, 1, 1)
        yi = F.batch_norm(xi, rm_ref, rv_ref, gamma, beta, training, 0.1, 1e-5)
        refs.append(torch.relu(yi.reshape(c, m).t() + res[gi]))
    assert _rel(out, torch.stack(refs)) < 1e-5
    assert _rel(rm_a, rm_ref) < 1e-5 and _rel(rv_a, rv_ref) < 1e-5


def test_maxpool_upsample(cuda):
    ops = _ops()
    gen = torch.Generator().manual_seed(9)
    x = torch.randn(2, 64, 9, 8, 7, generator=gen)
    out = ops.maxpool3d(x.permute(0, 2, 3, 4, 1).contiguous().to(cuda))
    ref = F.max_pool3d(x, 3, 2, 1).permute(0, 2, 3, 4, 1)
    assert torch.equal(out.cpu(), ref)
    coarse = torch.randn(2, 256, 4, 4, 4, generator=gen)
    lat = torch.randn(2, 256, 7, 8, 8, generator=gen)
    got = ops.upsample2_add(coarse.permute(0, 2, 3, 4, 1).contiguous().to(cuda),
                            lat.permute(0, 2, 3, 4, 1).contiguous().to(cuda))
    ref = (F.interpolate(coarse, scale_factor=2)[:, :, :7, :8, :8] + lat).permute(0, 2, 3, 4, 1)
    assert torch.equal(got.cpu(), ref)


def test_trilinear_gather(cuda):
    ops = _ops()
    gen = torch.Generator().manual_seed(11)
    X, Y, Z, c = 12, 10, 14, 256
    p1 = torch.randn(1, c, Z // 2, X // 2, Y // 2, generator=gen)
    store = torch.randn(X, Y, Z, 7, generator=gen)
    grid = store.permute(3, 2, 0, 1).unsqueeze(0)
    mask = torch.randperm(X * Y * Z, generator=gen)[:500].sort().values
    rows = ops.trilinear_gather(p1[0].permute(1, 2, 3, 0).contiguous().to(cuda),
                                store.to(cuda).permute(3, 2, 0, 1).unsqueeze(0), mask.to(cuda))
    up = F.interpolate(p1, size=(Z, X, Y), mode="trilinear", align_corners=True)
    ref_f = up.permute(0, 3, 4, 2, 1).reshape(1, -1, c)[0, mask]
    ref_xyz = grid[:, :3].permute(0, 3, 4, 2, 1).reshape(1, -1, 3)[0, mask]
    assert torch.equal(rows[:, :3].cpu(), ref_xyz)
    assert (rows[:, 4:].cpu() - ref_f).abs().max() < 1e-5


def test_downsample_bitexact(cuda):
    """R4 against the oracle's defined order / arithmetic: bit exact."""
    ops = _ops()
    from oracle.downsample import hierarchical_grid_subsample
    gen = torch.Generator().manual_seed(13)
    n_src, n_tgt = 5000, 4200
    pts = torch.rand(n_src + n_tgt, 3, generator=gen) * 2.4 - 1.2
    feats = torch.randn(n_src + n_tgt, 256, generator=gen)
    rows = torch.cat([pts, torch.zeros(n_src + n_tgt, 1), feats], dim=1)
    out, a, b = ops.hierarchical_downsample(rows.to(cuda), n_src, n_tgt, num_rounds=6)
    p_ref, f_ref, l_ref = hierarchical_grid_subsample(pts, feats, torch.tensor([n_src, n_tgt]), 6)
    assert [a, b] == l_ref.tolist()
    assert torch.equal(out[:, :3].cpu(), p_ref)
    assert torch.equal(out[:, 4:].cpu(), f_ref)


def test_pos_embed_and_layernorm(cuda):
    ops = _ops()
    from oracle.regtr import pos_embed_sine
    gen = torch.Generator().manual_seed(17)
    xyz = torch.rand(777, 3, generator=gen) * 3 - 1.5
    pe = ops.pos_embed_sine(xyz.to(cuda), 1.0)
    assert (pe.cpu() - pos_embed_sine(xyz)).abs().max() < 2e-5
    x = torch.randn(777, 256, generator=gen) * 2 + 0.3
    gamma, beta = torch.rand(256, generator=gen) + 0.5, torch.randn(256, generator=gen)
    out, (hi, lo) = ops.layernorm256(x.to(cuda), gamma.to(cuda), beta.to(cuda), add=pe, want_planes=True)
    ref = F.layer_norm(x, (256,), gamma, beta, 1e-5) + pos_embed_sine(xyz)
    assert (out.cpu() - ref).abs().max() < 2e-5
    assert _rel(hi.float() + lo.float(), ref) < 4e-6


@pytest.mark.parametrize("nq,nk", [(100, 100), (333, 517), (1500, 1400), (31, 5)])
def test_mha_core(cuda, nq, nk):
    ops = _ops()
    gen = torch.Generator().manual_seed(nq + nk)
    qkv_q = torch.randn(nq, 768, generator=gen)
    qkv_k = torch.randn(nk, 768, generator=gen)
    qd, kd = qkv_q.to(cuda), qkv_k.to(cuda)
    out = ops.mha_core(qd[:, :256], kd[:, 256:512], kd[:, 512:])
    q = qkv_q[:, :256].double().view(nq, 8, 32).transpose(0, 1)
    k = qkv_k[:, 256:512].double().view(nk, 8, 32).transpose(0, 1)
    v = qkv_k[:, 512:].double().view(nk, 8, 32).transpose(0, 1)
    att = torch.softmax(q @ k.transpose(1, 2) / math.sqrt(32), dim=-1)
    ref = (att @ v).transpose(0, 1).reshape(nq, 256)
    assert _rel(out, ref) < 1e-5


def test_decoder_tail_and_overlap(cuda):
    ops = _ops()
    gen = torch.Generator().manual_seed(23)
    nq, nk = 400, 517
    s = torch.randn(nq, 520, generator=gen) * 4
    xyz = torch.randn(nk, 3, generator=gen)
    out = ops.softmax_weighted_xyz(s.to(cuda), nk, xyz.to(cuda))
    ref = torch.softmax(s[:, :nk].double(), dim=-1) @ xyz.double()
    assert _rel(out, ref) < 1e-5
    feat = torch.randn(nq, 256, generator=gen)
    w, b = torch.randn(256, generator=gen) / 16, torch.randn(1, generator=gen)
    ov = ops.overlap_sigmoid(feat.to(cuda), w.to(cuda), b.to(cuda))
    assert _rel(ov, torch.sigmoid(feat.double() @ w.double() + b.double())) < 1e-5


def test_procrustes(cuda):
    ops = _ops()
    from oracle.regtr import compute_rigid_transform
    gen = torch.Generator().manual_seed(29)
    L, n = 6, 900
    a = torch.randn(L, n, 3, generator=gen)
    ang = 0.7
    R = torch.tensor([[math.cos(ang), -math.sin(ang), 0], [math.sin(ang), math.cos(ang), 0], [0, 0, 1.0]])
    t = torch.tensor([0.3, -0.2, 0.5])
    b = a @ R.t() + t + 0.01 * torch.randn(L, n, 3, generator=gen)
    w = torch.rand(L, n, generator=gen)
    out = ops.procrustes(a.to(cuda), b.to(cuda), w.to(cuda))
    ref = compute_rigid_transform(a, b, w)
    assert (out.cpu() - ref).abs().max() < 1e-5
    assert (out.cpu()[0, :, :3] - R).abs().max() < 5e-3
