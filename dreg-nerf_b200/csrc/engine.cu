// Whole-path engine: NeRFRegTr.forward (conerf/register/nerf_regtr.py:112-248) for one (src, tgt)
// pair without Python in the loop.  The two grids of a pair run through the FPN as g = 2 images of
// one launch sequence (BatchNorm statistics stay per grid, as in the reference's two B = 1 calls).
//
// Network topology restated from conerf/model/resnet3d.py:116-172 (ResNet-50, Bottleneck
// [3,4,6,3], conv1 5^3 s2, maxpool 3 s2) and conerf/model/feature_pyramid_net.py:39-108 (FPN v1),
// conerf/register/transformer.py:225-299 (pre-LN cross encoder), nerf_regtr.py:350-394 (decoder).
#include "engine.cuh"

using namespace drb;

namespace drb {

static ConvW make_conv(drb_engine* e, const std::string& name, int cout, int cin, int k, int stride,
                       bool bias) {
  ConvW w;
  w.cout = cout; w.cin = cin; w.k = k; w.stride = stride;
  w.p_w = e->add_param(name + ".weight", (long long)cout * cin * k * k * k, true);
  if (bias) w.p_b = e->add_param(name + ".bias", cout, true);
  const int taps = k * k * k;
  w.im2col = (stride != 1) || (cin % 64 != 0);
  if (w.im2col) {
    w.kpad = ((taps * cin + 63) / 64) * 64;
    w.hi = e->alloc<plane_t>((long long)cout * w.kpad);
    if (e->cfg.planes == 2) w.lo = e->alloc<plane_t>((long long)cout * w.kpad);
  } else {
    w.hi = e->alloc<plane_t>((long long)taps * cout * cin);
    if (e->cfg.planes == 2) w.lo = e->alloc<plane_t>((long long)taps * cout * cin);
  }
  return w;
}

static BnP make_bn(drb_engine* e, const std::string& name, int c) {
  BnP b;
  b.c = c;
  b.p_w = e->add_param(name + ".weight", c, true);
  b.p_b = e->add_param(name + ".bias", c, true);
  b.p_rm = e->add_param(name + ".running_mean", c, false);
  b.p_rv = e->add_param(name + ".running_var", c, false);
  b.scale = e->alloc<float>((long long)kG * c);
  b.shift = e->alloc<float>((long long)kG * c);
  b.mean = e->alloc<float>((long long)kG * c);
  b.rstd = e->alloc<float>((long long)kG * c);
  return b;
}

static Act make_act(drb_engine* e, int d, int h, int w, int c, bool f32, bool planes) {
  Act a;
  a.d = d; a.h = h; a.w = w; a.c = c;
  if (f32) a.f = e->alloc<float>(a.numel());
  if (planes) {
    a.hi = e->alloc<plane_t>(a.numel());
    if (e->cfg.planes == 2) a.lo = e->alloc<plane_t>(a.numel());
  }
  return a;
}

static inline int conv_out(int n, int k, int s, int p) { return (n + 2 * p - k) / s + 1; }

static int build(drb_engine* e) {
  const drb_engine_config& cfg = e->cfg;
  e->D = cfg.res_z; e->H = cfg.res_x; e->W = cfg.res_y;
  const std::string bb = "fpn3d.backbone_net";
  e->conv1 = make_conv(e, bb + ".conv1", 64, 4, 5, 2, false);
  e->bn1 = make_bn(e, bb + ".bn1", 64);
  const int nblk[4] = {3, 4, 6, 3};
  const int planes[4] = {64, 128, 256, 512};
  int inplanes = 64;
  for (int li = 0; li < 4; ++li) {
    for (int bi = 0; bi < nblk[li]; ++bi) {
      Block b;
      const std::string pn = bb + ".layer" + std::to_string(li + 1) + "." + std::to_string(bi);
      b.stride = (li > 0 && bi == 0) ? 2 : 1;
      const int pl = planes[li];
      b.conv1 = make_conv(e, pn + ".conv1", pl, inplanes, 1, 1, false);
      b.bn1 = make_bn(e, pn + ".bn1", pl);
      b.conv2 = make_conv(e, pn + ".conv2", pl, pl, 3, b.stride, false);
      b.bn2 = make_bn(e, pn + ".bn2", pl);
      b.conv3 = make_conv(e, pn + ".conv3", pl * 4, pl, 1, 1, false);
      b.bn3 = make_bn(e, pn + ".bn3", pl * 4);
      b.has_down = (bi == 0);
      if (b.has_down) {
        b.down = make_conv(e, pn + ".downsample.0", pl * 4, inplanes, 1, b.stride, false);
        b.bnd = make_bn(e, pn + ".downsample.1", pl * 4);
      }
      inplanes = pl * 4;
      e->blocks[li].push_back(b);
    }
  }
  const std::string fp = "fpn3d.feature_pyramid.";
  const int lat_cin[5] = {64, 256, 512, 1024, 2048};
  for (int i = 0; i < 5; ++i)
    e->pyr[i] = make_conv(e, fp + "pyramid_transformation_" + std::to_string(i + 1), 256, lat_cin[i],
                          i == 0 ? 3 : 1, 1, true);
  for (int i = 0; i < 4; ++i)
    e->ups[i] = make_conv(e, fp + "upsample_transform_" + std::to_string(i + 1), 256, 256, 3, 1, true);
  for (int l = 0; l < kLayers; ++l) {
    const std::string pn = "transformer_encoder.layers." + std::to_string(l);
    TLayer& t = e->tl[l];
    auto attn = [&](const std::string& an) {
      AttnW a;
      a.in_proj.cout = 768; a.in_proj.cin = 256; a.in_proj.k = 1; a.in_proj.stride = 1;
      a.in_proj.p_w = e->add_param(pn + "." + an + ".in_proj_weight", 768 * 256, true);
      a.in_proj.p_b = e->add_param(pn + "." + an + ".in_proj_bias", 768, true);
      a.in_proj.hi = e->alloc<plane_t>(768 * 256);
      if (e->cfg.planes == 2) a.in_proj.lo = e->alloc<plane_t>(768 * 256);
      a.out_proj = make_conv(e, pn + "." + an + ".out_proj", 256, 256, 1, 1, true);
      return a;
    };
    t.self_attn = attn("self_attn");
    t.cross_attn = attn("cross_attn");
    t.lin1 = make_conv(e, pn + ".linear1", 1024, 256, 1, 1, true);
    t.lin2 = make_conv(e, pn + ".linear2", 256, 1024, 1, 1, true);
    t.n1w = e->add_param(pn + ".norm1.weight", 256, true); t.n1b = e->add_param(pn + ".norm1.bias", 256, true);
    t.n2w = e->add_param(pn + ".norm2.weight", 256, true); t.n2b = e->add_param(pn + ".norm2.bias", 256, true);
    t.n3w = e->add_param(pn + ".norm3.weight", 256, true); t.n3b = e->add_param(pn + ".norm3.bias", 256, true);
  }
  e->fin_w = e->add_param("transformer_encoder.norm.weight", 256, true);
  e->fin_b = e->add_param("transformer_encoder.norm.bias", 256, true);
  e->q_proj = make_conv(e, "correspondence_decoder.q_proj", 256, 256, 1, 1, true);
  e->k_proj = make_conv(e, "correspondence_decoder.k_proj", 256, 256, 1, 1, true);
  e->conf_w = e->add_param("correspondence_decoder.conf_logits_decoder.weight", 256, true);
  e->conf_b = e->add_param("correspondence_decoder.conf_logits_decoder.bias", 1, true);

  // ---------------- activation buffers ----------------
  const int d1 = conv_out(e->D, 5, 2, 2), h1 = conv_out(e->H, 5, 2, 2), w1 = conv_out(e->W, 5, 2, 2);
  e->c1 = make_act(e, d1, h1, w1, 64, true, true);
  long long col_max = (long long)kG * e->c1.m() * e->conv1.kpad;
  long long raw_max = e->c1.numel();
  const int d2 = conv_out(d1, 3, 2, 1), h2 = conv_out(h1, 3, 2, 1), w2 = conv_out(w1, 3, 2, 1);
  e->x0 = make_act(e, d2, h2, w2, 64, true, true);
  int cd = d2, ch = h2, cw = w2;
  for (int li = 0; li < 4; ++li) {
    for (size_t bi = 0; bi < e->blocks[li].size(); ++bi) {
      const Block& b = e->blocks[li][bi];
      const int pl = planes[li];
      const int od = conv_out(cd, 3, b.stride, 1), oh = conv_out(ch, 3, b.stride, 1),
                ow = conv_out(cw, 3, b.stride, 1);
      // t1: conv1 output at the input resolution; t2 / out at the output resolution
      Act t1 = make_act(e, cd, ch, cw, pl, b.stride != 1, b.stride == 1);
      Act t2 = make_act(e, od, oh, ow, pl, false, true);
      Act out = make_act(e, od, oh, ow, pl * 4, true, true);
      e->tmp.push_back(t1); e->tmp.push_back(t2); e->tmp.push_back(out);
      if (t1.numel() > raw_max) raw_max = t1.numel();
      if (out.numel() > raw_max) raw_max = out.numel();
      if (b.conv2.im2col) {
        const long long n = (long long)kG * od * oh * ow * b.conv2.kpad;
        if (n > col_max) col_max = n;
      }
      if (b.has_down && b.down.im2col) {
        const long long n = (long long)kG * od * oh * ow * b.down.kpad;
        if (n > col_max) col_max = n;
      }
      cd = od; ch = oh; cw = ow;
    }
    e->c[li] = e->tmp.back();
  }
  e->col_elems = col_max;
  e->col_hi = e->alloc<plane_t>(col_max);
  if (cfg.planes == 2) e->col_lo = e->alloc<plane_t>(col_max);
  if ((long long)e->D * e->H * e->W * 4 > raw_max) raw_max = (long long)e->D * e->H * e->W * 4;   // rgba pack scratch
  e->raw_elems = raw_max;
  e->raw = e->alloc<float>(raw_max);
  e->raw2 = e->alloc<float>(raw_max);
  e->bn_accum = e->alloc<double>((long long)kG * 2048 * 2);
  e->splitk_ws_bytes = (size_t)24 << 20;     // <= 148 (tile, slice) pairs of 128 x 256 fp32
  e->splitk_ws = e->alloc<uint8_t>((long long)e->splitk_ws_bytes);
  // FPN: lat[i] / p[i] live at the resolution of c(i+1); c1 for i = 0
  const Act* feats[5] = {&e->c1, &e->c[0], &e->c[1], &e->c[2], &e->c[3]};
  for (int i = 0; i < 5; ++i) {
    const Act& f = *feats[i];
    e->p[i] = make_act(e, f.d, f.h, f.w, 256, true, false);
    if (i < 4) {
      e->lat[i] = make_act(e, f.d, f.h, f.w, 256, true, false);
      e->sum[i] = make_act(e, f.d, f.h, f.w, 256, false, true);
    }
  }
  {
    int box[4], tiles[4];
    drb_conv3d_tile_shape(kG, e->c1.d, e->c1.h, e->c1.w, box, tiles);
    const long long nt = (long long)tiles[0] * tiles[1] * tiles[2] * tiles[3];
    e->need = e->alloc<uint8_t>((long long)kG * e->c1.m());
    e->tiles_out = e->alloc<int>(nt);
    e->tiles_in = e->alloc<int>(nt);
    e->tiles_in2 = e->alloc<int>(nt);
    e->tile_counts = e->alloc<int>(3);
    e->tile_totals = e->alloc<unsigned long long>(3);
    cudaMemset(e->tile_totals, 0, 3 * sizeof(unsigned long long));
    // rows of tiles that are skipped keep whatever they held: start from zeros, not from garbage
    if (e->lat[0].f) cudaMemset(e->lat[0].f, 0, sizeof(float) * e->lat[0].numel());
    if (e->p[0].f) cudaMemset(e->p[0].f, 0, sizeof(float) * e->p[0].numel());
    if (e->sum[0].hi) cudaMemset(e->sum[0].hi, 0, sizeof(plane_t) * e->sum[0].numel());
    if (e->sum[0].lo) cudaMemset(e->sum[0].lo, 0, sizeof(plane_t) * e->sum[0].numel());
  }
  e->rows = e->alloc<float>((long long)2 * cfg.max_mask * kRowLd);
  e->rows_ds = e->alloc<float>((long long)2 * cfg.max_mask * kRowLd);
  e->ds_ws_bytes = drb_downsample_workspace_bytes(2 * cfg.max_mask, kRowLd);
  e->ds_ws = e->alloc<uint8_t>((long long)e->ds_ws_bytes);
  // every packed weight, in one table (the vectors above are final: the pointers stay valid)
  e->conv1.need_dgrad = false;
  e->all_convs.push_back(&e->conv1);
  for (int li = 0; li < 4; ++li)
    for (Block& b : e->blocks[li]) {
      e->all_convs.push_back(&b.conv1); e->all_convs.push_back(&b.conv2); e->all_convs.push_back(&b.conv3);
      if (b.has_down) e->all_convs.push_back(&b.down);
    }
  for (int i = 0; i < 5; ++i) e->all_convs.push_back(&e->pyr[i]);
  for (int i = 0; i < 4; ++i) e->all_convs.push_back(&e->ups[i]);
  for (int l = 0; l < kLayers; ++l) {
    TLayer& t = e->tl[l];
    e->all_convs.push_back(&t.self_attn.in_proj); e->all_convs.push_back(&t.self_attn.out_proj);
    e->all_convs.push_back(&t.cross_attn.in_proj); e->all_convs.push_back(&t.cross_attn.out_proj);
    e->all_convs.push_back(&t.lin1); e->all_convs.push_back(&t.lin2);
  }
  e->all_convs.push_back(&e->q_proj); e->all_convs.push_back(&e->k_proj);
  const int nw = (int)e->all_convs.size();
  e->n_slots = nw + 64;                                   // one per weight + a ring for gradient tensors
  e->slots = e->alloc<float>(2LL * e->n_slots);
  if (e->slots) cudaMemset(e->slots, 0, sizeof(float) * 2 * e->n_slots);
  e->d_pack = e->alloc<PackDesc>(nw);
  for (int i = 0; i < nw; ++i) {
    e->all_convs[i]->pack_index = i;
    e->all_convs[i]->slot = e->slots ? e->slots + 2 * i : nullptr;
  }
  return e->fail.empty() ? 0 : DRB_ENOMEM;
}

// --------------------------------------------------------------------------------------------

int engine_list_id(const drb_engine* e, const int* tile_list) {
  if (tile_list == nullptr) return -1;
  return tile_list == e->tiles_out ? 0 : (tile_list == e->tiles_in ? 1 : 2);
}

int engine_run_igemm(drb_engine* e, const ConvW& w, const plane_t* x_hi, const plane_t* x_lo, int g, int d,
                     int h, int wd, int cin, int k, const float* bias, const float* residual,
                     int relu, float scale, float* out, plane_t* out_hi, plane_t* out_lo, long long ld,
                     cudaStream_t s, const plane_t* w_hi, const plane_t* w_lo, int cout_override,
                     const int* tile_list, const int* tile_count, const float* scale_dev0,
                     const float* scale_dev1, const int* res_dims) {
  drb_conv3d_desc cd;
  memset(&cd, 0, sizeof(cd));
  cd.g = g; cd.d = d; cd.h = h; cd.w = wd;
  cd.cin = cin; cd.cout = cout_override ? cout_override : w.cout;
  cd.kd = cd.kh = cd.kw = k;
  cd.planes = e->cfg.planes;
  cd.relu = relu; cd.out_scale = scale;
  cd.acc_scale = 1.f;
  cd.x_hi = x_hi; cd.x_lo = x_lo;
  cd.w_hi = w_hi ? w_hi : w.hi; cd.w_lo = w_lo ? w_lo : w.lo;
  cd.bias = bias; cd.residual = residual;
  cd.out = out; cd.out_hi = out_hi; cd.out_lo = out_lo;
  cd.ld_out = ld;
  cd.tile_list = tile_list; cd.tile_count = tile_count;
  // the weight pre-scale lives on the device (drb_engine_commit_params never synchronises)
  cd.acc_scale_dev[0] = scale_dev0;
  cd.acc_scale_dev[1] = scale_dev1;
  cd.splitk_ws = e->splitk_ws;
  cd.splitk_ws_bytes = e->splitk_ws_bytes;
  if (res_dims) { cd.res_d = res_dims[0]; cd.res_h = res_dims[1]; cd.res_w = res_dims[2]; }
  e->launches += 1;
  if (!e->profile) return drb_conv3d_igemm(&cd, s);
  drb_engine::ProfRec r;
  cudaEventCreate(&r.a);
  cudaEventCreate(&r.b);
  r.flops = 2.0 * (double)g * d * h * wd * (double)cd.cout * (double)cin * k * k * k;
  r.list_id = engine_list_id(e, tile_list);
  r.flops_per_tile = 2.0 * 128.0 * (double)cd.cout * (double)cin * k * k * k;   // executed work of one listed tile
  r.m = g * d * h * wd; r.cin = cin; r.cout = cd.cout; r.k = k;
  cudaEventRecord(r.a, s);
  const int rc = drb_conv3d_igemm(&cd, s);
  cudaEventRecord(r.b, s);
  e->prof.push_back(r);
  return rc;
}

// forward GEMM with the layer's own packed weights (their inverse pre-scale is slot + 1)
static int run_igemm(drb_engine* e, const ConvW& w, const plane_t* x_hi, const plane_t* x_lo, int g, int d,
                     int h, int wd, int cin, int k, const float* bias, const float* residual,
                     int relu, float scale, float* out, plane_t* out_hi, plane_t* out_lo, long long ld,
                     cudaStream_t s, const plane_t* w_hi = nullptr, const plane_t* w_lo = nullptr,
                     int cout_override = 0, const int* tile_list = nullptr, const int* tile_count = nullptr,
                     const int* res_dims = nullptr) {
  const float* wscale = (w_hi == nullptr && e->cfg.planes == 2 && w.slot) ? w.slot + 1 : nullptr;
  return engine_run_igemm(e, w, x_hi, x_lo, g, d, h, wd, cin, k, bias, residual, relu, scale, out, out_hi, out_lo, ld,
                          s, w_hi, w_lo, cout_override, tile_list, tile_count, wscale, nullptr, res_dims);
}


// im2col of `in` for the strided / narrow convolution `w` into the shared column planes
int engine_im2col(drb_engine* e, const ConvW& w, const Act& in, cudaStream_t s) {
  drb_im2col_desc d;
  memset(&d, 0, sizeof(d));
  d.x = in.f;
  d.sc = 1; d.sw = in.c; d.sh = (long long)in.w * in.c; d.sd = (long long)in.h * in.w * in.c;
  d.sg = (long long)in.d * in.h * in.w * in.c;
  d.g = kG; d.c = in.c; d.d = in.d; d.h = in.h; d.w = in.w;
  d.k = w.k; d.stride = w.stride; d.pad = w.k / 2; d.kpad = w.kpad;
  e->launches += 1;
  return drb_im2col(&d, e->col_hi, e->col_lo, s);
}

// conv (stride 1 via TMA implicit GEMM, otherwise im2col + 1x1x1 GEMM) into raw fp32 [g][m][cout].
// Every caller feeds a BatchNorm.  The GEMM epilogue CAN produce the per-channel sums
// (drb_conv3d_desc.bn_accum, parity-tested), but its fp64 atomics were measured slower than the separate
// HBM-bound pass on B200 (+1.2 ms vs -0.85 ms per pair, round 1), so the engine keeps the separate pass.
static int conv_any(drb_engine* e, const ConvW& w, const Act& in, int od, int oh, int ow, float* out,
                    cudaStream_t s) {
  if (!w.im2col)
    return run_igemm(e, w, in.hi, in.lo, kG, in.d, in.h, in.w, w.cin, w.k, P(e, w.p_b), nullptr, 0,
                     1.f, out, nullptr, nullptr, 0, s);
  DRB_TRY(engine_im2col(e, w, in, s));
  return run_igemm(e, w, e->col_hi, e->col_lo, kG, od, oh, ow, w.kpad, 1, P(e, w.p_b), nullptr, 0, 1.f,
                   out, nullptr, nullptr, 0, s);
}

// BatchNorm (+ residual, ReLU) of raw [g][m][c] into out (fp32 and/or planes)
static int bn_apply(drb_engine* e, const BnP& b, const float* rawp, long long m, const float* residual,
                    int relu, float* out, plane_t* out_hi, plane_t* out_lo, cudaStream_t s) {
  const int training = e->cfg.training_bn;
  if (e->bn_small && drb_bn_small_supported(m, b.c)) {
    // deep stages: statistics + running update + apply (+ the backward's saved statistics) in one launch
    const bool upd_s = !training || e->update_running;
    e->launches += 1;
    return drb_bn_small(rawp, kG, m, b.c, P(e, b.p_w), P(e, b.p_b), upd_s ? P(e, b.p_rm) : nullptr,
                        upd_s ? P(e, b.p_rv) : nullptr, training, 0.1f, 1e-5f, residual, relu, out, out_hi, out_lo,
                        e->grad_mode ? b.mean : nullptr, e->grad_mode ? b.rstd : nullptr,
                        e->grad_mode ? b.scale : nullptr, e->grad_mode ? b.shift : nullptr, s);
  }
  if (training) {
    e->launches += 2;
    DRB_TRY(drb_bn_stats(rawp, kG, m, b.c, e->bn_accum, s));
  }
  if (e->grad_mode) {   // what the backward pass needs: the statistics the apply below is about to use
    e->launches += 1;
    DRB_TRY(drb_bn_save_stats(e->bn_accum, kG, m, b.c, P(e, b.p_w), P(e, b.p_b), P(e, b.p_rm), P(e, b.p_rv), training,
                              1e-5f, b.mean, b.rstd, b.scale, b.shift, s));
  }
  e->launches += 1;
  // a secondary engine of a stream pipeline (drb_engine_set_update_running(e, 0)) normalises with its own batch
  // statistics but leaves the shared running buffers to the primary engine (the replicas of a DDP job do the same)
  const bool upd = !training || e->update_running;
  return drb_bn_apply(rawp, e->bn_accum, kG, m, b.c, P(e, b.p_w), P(e, b.p_b), upd ? P(e, b.p_rm) : nullptr,
                      upd ? P(e, b.p_rv) : nullptr, training, 0.1f, 1e-5f, residual, relu, out, out_hi, out_lo, s);
}

// conv1 (5^3 s2, Cin 4 = rgba channels 3..6 of the [1,7,Z,X,Y] grid): im2col of both grids into the shared
// column planes
int engine_stem_im2col(drb_engine* e, const drb_pair_io* io, cudaStream_t s) {
  const ConvW& w = e->conv1;
  const long long per_grid = e->c1.m() * w.kpad;
  for (int g = 0; g < kG; ++g) {
    const float* base = g == 0 ? io->src_grid : io->tgt_grid;
    const long long sc = g == 0 ? io->s_ch : io->t_ch;
    e->launches += 2;
    DRB_TRY(drb_im2col_stem(base + 3 * sc, sc, g == 0 ? io->s_z : io->t_z, g == 0 ? io->s_x : io->t_x,
                            g == 0 ? io->s_y : io->t_y, e->D, e->H, e->W, e->raw2,
                            e->col_hi + g * per_grid, off(e->col_lo, g * per_grid), s));
  }
  return 0;
}

static int run_fpn(drb_engine* e, const drb_pair_io* io, cudaStream_t s) {
  // in grad mode every BatchNorm keeps the convolution output it normalised; otherwise one scratch is reused
  auto rawbuf = [&](const BnP& b, float* shared) { return e->grad_mode ? b.raw_keep : shared; };
  {
    const ConvW& w = e->conv1;
    DRB_TRY(engine_stem_im2col(e, io, s));
    float* r0 = rawbuf(e->bn1, e->raw);
    DRB_TRY(run_igemm(e, w, e->col_hi, e->col_lo, kG, e->c1.d, e->c1.h, e->c1.w, w.kpad, 1, nullptr,
                      nullptr, 0, 1.f, r0, nullptr, nullptr, 0, s));
    DRB_TRY(bn_apply(e, e->bn1, r0, e->c1.m(), nullptr, 1, e->c1.f, e->c1.hi, e->c1.lo, s));
    e->launches += 1;
    DRB_TRY(drb_maxpool3d(e->c1.f, kG, e->c1.d, e->c1.h, e->c1.w, 64, e->x0.f, e->x0.hi, e->x0.lo, s));
  }
  // ---- bottleneck stacks ----
  const Act* x = &e->x0;
  size_t ti = 0;
  for (int li = 0; li < 4; ++li) {
    for (size_t bi = 0; bi < e->blocks[li].size(); ++bi) {
      const Block& b = e->blocks[li][bi];
      const Act& t1 = e->tmp[ti]; const Act& t2 = e->tmp[ti + 1]; const Act& out = e->tmp[ti + 2];
      ti += 3;
      float* r1 = rawbuf(b.bn1, e->raw);
      DRB_TRY(conv_any(e, b.conv1, *x, x->d, x->h, x->w, r1, s));
      DRB_TRY(bn_apply(e, b.bn1, r1, t1.m(), nullptr, 1, t1.f, t1.hi, t1.lo, s));
      float* r2 = rawbuf(b.bn2, e->raw);
      DRB_TRY(conv_any(e, b.conv2, t1, t2.d, t2.h, t2.w, r2, s));
      DRB_TRY(bn_apply(e, b.bn2, r2, t2.m(), nullptr, 1, nullptr, t2.hi, t2.lo, s));
      const float* res = x->f;
      if (b.has_down) {
        float* rd = rawbuf(b.bnd, e->raw2);
        DRB_TRY(conv_any(e, b.down, *x, out.d, out.h, out.w, rd, s));
        DRB_TRY(bn_apply(e, b.bnd, rd, out.m(), nullptr, 0, e->raw2, nullptr, nullptr, s));
        res = e->raw2;
      }
      float* r3 = rawbuf(b.bn3, e->raw);
      DRB_TRY(conv_any(e, b.conv3, t2, out.d, out.h, out.w, r3, s));
      DRB_TRY(bn_apply(e, b.bn3, r3, out.m(), res, 1, out.f, out.hi, out.lo, s));
      x = &out;
    }
  }
  // ---- feature pyramid (top-down) ----
  const Act* feats[5] = {&e->c1, &e->c[0], &e->c[1], &e->c[2], &e->c[3]};
  {
    const Act& f = *feats[4];
    DRB_TRY(run_igemm(e, e->pyr[4], f.hi, f.lo, kG, f.d, f.h, f.w, f.c, 1, P(e, e->pyr[4].p_b), nullptr, 0,
                      1.f, e->p[4].f, nullptr, nullptr, 0, s));
  }
  for (int i = 3; i >= 0; --i) {
    const Act& f = *feats[i];
    const bool sparse = (i == 0) && e->sparse_fpn;
    const Act& top = e->p[i + 1];
    // The two large levels: the top-down merge up(top) + lateral (feature_pyramid_net.py:58-61) happens in the lateral
    // convolution's epilogue (the residual is the coarser level, read with nearest x2 up-sampling) and goes straight
    // to the planes the smoothing convolution reads - no fp32 lateral tensor, no separate pass (0.54 GB written +
    // 0.54 GB read back at level 1).  The small levels keep the split-K lateral + separate merge.
    if (e->fuse_topdown && i <= 1) {
      const int rd[3] = {top.d, top.h, top.w};
      DRB_TRY(run_igemm(e, e->pyr[i], f.hi, f.lo, kG, f.d, f.h, f.w, f.c, e->pyr[i].k, P(e, e->pyr[i].p_b),
                        top.f, 0, 1.f, nullptr, e->sum[i].hi, e->sum[i].lo, 0, s, nullptr, nullptr, 0,
                        sparse ? e->tiles_in : nullptr, sparse ? e->tile_counts + 1 : nullptr, rd));
    } else {
      DRB_TRY(run_igemm(e, e->pyr[i], f.hi, f.lo, kG, f.d, f.h, f.w, f.c, e->pyr[i].k, P(e, e->pyr[i].p_b),
                        nullptr, 0, 1.f, e->lat[i].f, nullptr, nullptr, 0, s, nullptr, nullptr, 0,
                        sparse ? e->tiles_in : nullptr, sparse ? e->tile_counts + 1 : nullptr));
      e->launches += 1;
      DRB_TRY(drb_upsample2_add(top.f, top.d, top.h, top.w, e->lat[i].f, kG, f.d, f.h, f.w, 256, nullptr,
                                e->sum[i].hi, e->sum[i].lo, s));
    }
    DRB_TRY(run_igemm(e, e->ups[i], e->sum[i].hi, e->sum[i].lo, kG, f.d, f.h, f.w, 256, 3,
                      P(e, e->ups[i].p_b), nullptr, 0, 1.f, e->p[i].f, nullptr, nullptr, 0, s, nullptr, nullptr, 0,
                      sparse ? e->tiles_out : nullptr, sparse ? e->tile_counts : nullptr));
  }
  return 0;
}

__global__ void unpack_rows_kernel(const float* __restrict__ rows, int n, float* __restrict__ kp,
                                   float* __restrict__ x) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)n * 64) return;
  const int row = (int)(i >> 6), c4 = (int)(i & 63) * 4;
  const float* r = rows + (long long)row * kRowLd;
  *(float4*)(x + (long long)row * kD + c4) = *(const float4*)(r + 4 + c4);
  if (c4 == 0) { kp[row * 3] = r[0]; kp[row * 3 + 1] = r[1]; kp[row * 3 + 2] = r[2]; }
}

int engine_ensure_tokens(drb_engine* e, int m) {
  if (m <= e->tok_cap && (!e->grad_mode || e->tok_grad)) return 0;
  for (void* p : e->tok_allocs) cudaFree(p);
  e->tok_allocs.clear();
  if (m < e->tok_cap) m = e->tok_cap;
  // the down-sampler stops at max_tokens points for the pair: allocate for that bound once instead of growing
  // pair by pair (cudaFree / cudaMalloc wait for the whole device - every stream of a pipeline would stall)
  if (m < e->max_tokens && e->max_tokens <= 32768) m = e->max_tokens;
  const long long cap = ((m + 127) / 128) * 128 + 128;
  bool ok = true;
  auto A = [&](long long bytes) -> void* {
    void* p = nullptr;
    if (cudaMalloc(&p, (size_t)bytes) != cudaSuccess) { ok = false; return nullptr; }
    e->tok_allocs.push_back(p);
    return p;
  };
  const long long cap_ld = ((cap + 7) / 8) * 8;
  e->x = (float*)A(cap * kD * 4); e->pos = (float*)A(cap * kD * 4);
  e->qkv = (float*)A(cap * 768 * 4); e->kp_xyz = (float*)A(cap * 3 * 4 + 64);
  e->sbuf = (float*)A(cap * cap_ld * 4);
  const bool pair = e->cfg.planes == 2;
  auto AP = [&](long long elems) -> plane_t* { return (plane_t*)A(elems * 2); };
  auto AL = [&](long long elems) -> plane_t* { return pair ? (plane_t*)A(elems * 2) : nullptr; };
  e->xn_hi = AP(cap * kD); e->xn_lo = AL(cap * kD);
  e->att_hi = AP(cap * kD); e->att_lo = AL(cap * kD);
  e->ffn_hi = AP(cap * 1024); e->ffn_lo = AL(cap * 1024);
  e->dec_hi = AP(kLayers * cap * kD); e->dec_lo = AL(kLayers * cap * kD);
  e->qp_hi = AP(kLayers * cap * kD); e->qp_lo = AL(kLayers * cap * kD);
  e->kp_hi = AP(kLayers * cap * kD); e->kp_lo = AL(kLayers * cap * kD);
  e->att_ws_bytes = drb_mha_tc_workspace_bytes((int)cap, 8, e->cfg.planes);
  e->att_ws = A((long long)e->att_ws_bytes);
  e->tok_grad = e->grad_mode;
  if (e->grad_mode) {
    // the training graph: per-layer copies of everything the backward pass re-reads, plus its scratch
    for (int l = 0; l <= kLayers; ++l) e->xs[l] = (float*)A(cap * kD * 4);
    for (int l = 0; l < kLayers; ++l) {
      TSave& t = e->ts[l];
      t.x1 = (float*)A(cap * kD * 4); t.x2 = (float*)A(cap * kD * 4);
      t.xn1_hi = AP(cap * kD); t.xn1_lo = AL(cap * kD);
      t.xn2_hi = AP(cap * kD); t.xn2_lo = AL(cap * kD);
      t.xn3_hi = AP(cap * kD); t.xn3_lo = AL(cap * kD);
      t.qkv_s = (float*)A(cap * 768 * 4); t.qkv_c = (float*)A(cap * 768 * 4);
      t.att_s_hi = AP(cap * kD); t.att_s_lo = AL(cap * kD);
      t.att_c_hi = AP(cap * kD); t.att_c_lo = AL(cap * kD);
      t.ffn_hi = AP(cap * 1024); t.ffn_lo = AL(cap * 1024);
    }
    e->qf = (float*)A(kLayers * cap * kD * 4); e->kf = (float*)A(kLayers * cap * kD * 4);
    e->t_dx = (float*)A(cap * kD * 4); e->t_dy = (float*)A(cap * kD * 4);
    e->t_dh = (float*)A(cap * 1024 * 4); e->t_dqkv = (float*)A(cap * 768 * 4);
    e->t_datt = (float*)A(cap * kD * 4); e->t_dxn = (float*)A(cap * kD * 4);
    e->t_ddec = (float*)A(kLayers * cap * kD * 4);
    e->t_dqf = (float*)A(kLayers * cap * kD * 4); e->t_dkf = (float*)A(kLayers * cap * kD * 4);
    e->t_dcorr = (float*)A(kLayers * cap * 3 * 4 + 64); e->t_dov = (float*)A(kLayers * cap * 4 + 64);
    e->t_ph = AP(kLayers * cap * 1024 / 4 + cap * 1024); e->t_pl = AL(kLayers * cap * 1024 / 4 + cap * 1024);
    e->mha_ws_bytes = drb_mha_backward_workspace_bytes((int)cap, (int)cap, 8);
    e->mha_ws = A((long long)e->mha_ws_bytes);
  } else {
    for (int l = 0; l <= kLayers; ++l) e->xs[l] = e->x;
    for (int l = 0; l < kLayers; ++l) {
      TSave& t = e->ts[l];
      t.x1 = t.x2 = e->x;
      t.xn1_hi = t.xn2_hi = t.xn3_hi = e->xn_hi; t.xn1_lo = t.xn2_lo = t.xn3_lo = e->xn_lo;
      t.qkv_s = t.qkv_c = e->qkv;
      t.att_s_hi = t.att_c_hi = e->att_hi; t.att_s_lo = t.att_c_lo = e->att_lo;
      t.ffn_hi = e->ffn_hi; t.ffn_lo = e->ffn_lo;
    }
    e->qf = e->kf = nullptr;
  }
  if (!ok) {
    set_error("drb_engine: out of device memory for %d tokens", m);
    e->tok_cap = 0;
    return DRB_ENOMEM;
  }
  e->tok_cap = (int)cap;
  return 0;
}

// ------------------------------------------------------------------------------------------
// Multi-tensor weight packing: one absmax launch + one pack launch for all ~150 weights, pre-scales kept
// on the device (no host synchronisation; a training step repacks after every optimiser update).
//   forward layout   [tap][cout][cin]            (im2col: [cout][kpad], k = tap * cin + c)
//   data-grad layout [taps-1-tap][cin][cout]     (im2col: [kpad][cout])  - the same implicit GEMM then
//                    computes dX = dY (*) flip(W)^T, respectively dcol = dY W.
// ------------------------------------------------------------------------------------------
__global__ void pack_absmax_kernel(const PackDesc* __restrict__ descs) {
  const PackDesc d = descs[blockIdx.y];
  const long long n = (long long)d.cout * d.cin * d.taps;
  float m = 0.f;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) m = fmaxf(m, fabsf(d.w[i]));
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax((unsigned int*)d.slot, __float_as_uint(m));
}

__global__ void pack_all_kernel(const PackDesc* __restrict__ descs, int pair) {
  const PackDesc d = descs[blockIdx.y];
  float scale = 1.f;
  if (pair) {
    const float mx = __uint_as_float(*(const unsigned int*)d.slot);
    if (mx > 0.f && isfinite(mx)) {
      int ex = 0;
      frexpf(mx, &ex);
      scale = ldexpf(1.f, 9 - ex);          // max |w| * scale in [256, 512): fp16-normal hi AND lo
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) d.slot[1] = 1.f / scale;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const int cin = d.cin, cout = d.cout, taps = d.taps, kpad = d.kpad;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < d.fwd_elems; i += stride) {
    float v = 0.f;
    if (kpad > 0) {
      const int k = (int)(i % kpad), o = (int)(i / kpad);
      if (k < taps * cin) v = d.w[((long long)o * cin + k % cin) * taps + k / cin];
    } else {
      const int c = (int)(i % cin), o = (int)((i / cin) % cout), t = (int)(i / ((long long)cin * cout));
      v = d.w[((long long)o * cin + c) * taps + t];
    }
    plane_t h, l;
    split16(v * scale, pair != 0, h, l);
    d.hi[i] = h;
    if (pair) d.lo[i] = l;
  }
  if (!d.thi) return;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < d.bwd_elems; i += stride) {
    float v = 0.f;
    const int o = (int)(i % cout);
    if (kpad > 0) {
      const int k = (int)(i / cout);
      if (k < taps * cin) v = d.w[((long long)o * cin + k % cin) * taps + k / cin];
    } else {
      const int c = (int)((i / cout) % cin), t = (int)(i / ((long long)cin * cout));
      v = d.w[((long long)o * cin + c) * taps + (taps - 1 - t)];
    }
    plane_t h, l;
    split16(v * scale, pair != 0, h, l);
    d.thi[i] = h;
    if (pair) d.tlo[i] = l;
  }
}

}  // namespace drb

// ================================================================================================
extern "C" int drb_engine_create(const drb_engine_config* cfg, drb_engine** out) {
  DRB_REQUIRE(cfg && out, "drb_engine_create: null argument");
  DRB_REQUIRE(cfg->res_x >= 8 && cfg->res_y >= 8 && cfg->res_z >= 8, "drb_engine_create: resolution < 8");
  DRB_REQUIRE(cfg->planes == 1 || cfg->planes == 2, "drb_engine_create: planes must be 1 or 2");
  DRB_REQUIRE(cfg->max_mask > 0, "drb_engine_create: max_mask must be positive");
  drb_engine* e = new drb_engine();
  e->cfg = *cfg;
  if (const char* env = getenv("DRB_BN_SMALL")) e->bn_small = atoi(env) != 0;
  if (const char* env = getenv("DRB_FUSE_TOPDOWN")) e->fuse_topdown = atoi(env) != 0;
  int rc = build(e);
  if (rc) {
    set_error("drb_engine_create: %s", e->fail.c_str());
    drb_engine_destroy(e);
    return rc;
  }
  *out = e;
  return 0;
}

extern "C" void drb_engine_destroy(drb_engine* e) {
  if (!e) return;
  for (void* p : e->allocs) cudaFree(p);
  for (void* p : e->tok_allocs) cudaFree(p);
  delete e;
}

extern "C" int drb_engine_num_params(const drb_engine* e) { return e ? (int)e->params.size() : 0; }
extern "C" const char* drb_engine_param_name(const drb_engine* e, int i) {
  return (e && i >= 0 && i < (int)e->params.size()) ? e->params[i].name.c_str() : nullptr;
}
extern "C" long long drb_engine_param_numel(const drb_engine* e, int i) {
  return (e && i >= 0 && i < (int)e->params.size()) ? e->params[i].numel : -1;
}
extern "C" int drb_engine_bind_param(drb_engine* e, int i, float* device_ptr) {
  DRB_REQUIRE(e && i >= 0 && i < (int)e->params.size() && device_ptr, "drb_engine_bind_param: bad arguments");
  e->params[i].ptr = device_ptr;
  e->committed = false;
  return 0;
}
extern "C" int drb_engine_set_training(drb_engine* e, int training_bn) {
  DRB_REQUIRE(e, "drb_engine_set_training: null engine");
  e->cfg.training_bn = training_bn ? 1 : 0;
  return 0;
}
extern "C" long long drb_engine_launch_count(const drb_engine* e) { return e ? e->launches : 0; }

extern "C" int drb_engine_set_sparse_fpn(drb_engine* e, int on) {
  DRB_REQUIRE(e, "drb_engine_set_sparse_fpn: null engine");
  e->sparse_fpn = on != 0;
  return 0;
}

extern "C" int drb_engine_set_update_running(drb_engine* e, int on) {
  DRB_REQUIRE(e, "drb_engine_set_update_running: null engine");
  e->update_running = on != 0;
  return 0;
}

extern "C" int drb_engine_set_profile(drb_engine* e, int on) {
  DRB_REQUIRE(e, "drb_engine_set_profile: null engine");
  e->profile = on != 0;
  return 0;
}

extern "C" int drb_engine_profile_read(drb_engine* e, double* igemm_ms, double* igemm_flops, long long* n) {
  DRB_REQUIRE(e && igemm_ms && igemm_flops && n, "drb_engine_profile_read: null argument");
  double ms = 0.0, fl = 0.0;
  // output-sparse launches: executed FLOPs = FLOPs of one tile x the mean length of the launch's tile list
  unsigned long long totals[3] = {0, 0, 0};
  DRB_CUDA_OK(cudaDeviceSynchronize());
  if (e->tile_totals) {
    DRB_CUDA_OK(cudaMemcpy(totals, e->tile_totals, sizeof(totals), cudaMemcpyDeviceToHost));
    DRB_CUDA_OK(cudaMemset(e->tile_totals, 0, sizeof(totals)));
  }
  const double forwards = e->prof_forwards > 0 ? (double)e->prof_forwards : 1.0;
  const bool dump = getenv("DRB_PROFILE_DUMP") != nullptr;
  for (auto& r : e->prof) {
    DRB_CUDA_OK(cudaEventSynchronize(r.b));
    float t = 0.f;
    DRB_CUDA_OK(cudaEventElapsedTime(&t, r.a, r.b));
    ms += t;
    const double f_exec = r.list_id >= 0 ? r.flops_per_tile * ((double)totals[r.list_id] / forwards) : r.flops;
    fl += f_exec;
    if (dump)     // DRB_PROFILE_DUMP=1: one line per launch (development aid: which GEMM shapes run below the average)
      fprintf(stderr, "igemm M=%d Cin=%d Cout=%d k=%d%s: %.1f us, %.1f TFLOP/s executed\n", r.m, r.cin, r.cout, r.k,
              r.list_id >= 0 ? " (tile list)" : "", t * 1e3, t > 0.f ? f_exec / (t * 1e-3) / 1e12 : 0.0);
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  *igemm_ms = ms; *igemm_flops = fl; *n = (long long)e->prof.size();
  e->prof.clear();
  e->prof_forwards = 0;
  return 0;
}

extern "C" int drb_engine_commit_params(drb_engine* e, cudaStream_t s) {
  DRB_REQUIRE(e, "drb_engine_commit_params: null engine");
  for (const Param& p : e->params)
    DRB_REQUIRE(p.ptr != nullptr, "drb_engine_commit_params: parameter %s is not bound", p.name.c_str());
  if (e->grad_mode) DRB_TRY(engine_ensure_grad_buffers(e));
  const bool pair = e->cfg.planes == 2;
  const int nw = (int)e->all_convs.size();
  std::vector<PackDesc> table((size_t)nw);
  for (int i = 0; i < nw; ++i) {
    const ConvW& w = *e->all_convs[i];
    PackDesc& d = table[(size_t)i];
    const int taps = w.k * w.k * w.k;
    d.w = P(e, w.p_w);
    d.cout = w.cout; d.cin = w.cin; d.taps = taps; d.kpad = w.im2col ? w.kpad : 0;
    d.hi = w.hi; d.lo = pair ? w.lo : nullptr;
    d.thi = w.thi; d.tlo = pair ? w.tlo : nullptr;
    d.slot = w.slot;
    d.fwd_elems = w.im2col ? (long long)w.cout * w.kpad : (long long)taps * w.cout * w.cin;
    d.bwd_elems = d.fwd_elems;
  }
  DRB_CUDA_OK(cudaMemcpyAsync(e->d_pack, table.data(), sizeof(PackDesc) * (size_t)nw, cudaMemcpyHostToDevice, s));
  DRB_CUDA_OK(cudaMemsetAsync(e->slots, 0, sizeof(float) * 2 * (size_t)nw, s));
  e->launches += 2;
  if (pair) {
    pack_absmax_kernel<<<dim3(16, (unsigned)nw), 256, 0, s>>>(e->d_pack);
    DRB_LAUNCH_OK();
  }
  pack_all_kernel<<<dim3(64, (unsigned)nw), 256, 0, s>>>(e->d_pack, pair ? 1 : 0);
  DRB_LAUNCH_OK();
  e->committed = true;
  return 0;
}

extern "C" int drb_engine_set_grad_mode(drb_engine* e, int on) {
  DRB_REQUIRE(e, "drb_engine_set_grad_mode: null engine");
  const bool want = on != 0;
  if (want && !e->grad_allocated) e->committed = false;   // the data-gradient weight planes must be packed
  e->grad_mode = want;
  if (!want) e->graph_valid = false;
  return 0;
}
extern "C" int drb_engine_param_trainable(const drb_engine* e, int i) {
  return (e && i >= 0 && i < (int)e->params.size() && e->params[i].trainable) ? 1 : 0;
}
extern "C" int drb_engine_bind_grad(drb_engine* e, int i, float* device_ptr) {
  DRB_REQUIRE(e && i >= 0 && i < (int)e->params.size(), "drb_engine_bind_grad: bad arguments");
  e->params[i].grad = device_ptr;
  return 0;
}
extern "C" int drb_engine_set_tc_attention(drb_engine* e, int on) {
  DRB_REQUIRE(e, "drb_engine_set_tc_attention: null engine");
  e->tc_attention = on != 0;
  return 0;
}
extern "C" int drb_engine_set_max_tokens(drb_engine* e, int max_total) {
  DRB_REQUIRE(e && max_total > 0, "drb_engine_set_max_tokens: bad arguments");
  e->max_tokens = max_total;
  return 0;
}

extern "C" int drb_engine_encode(drb_engine* e, const drb_pair_io* io, int* host_n_src, int* host_n_tgt,
                                 cudaStream_t s) {
  DRB_REQUIRE(e && io && host_n_src && host_n_tgt, "drb_engine_encode: null argument");
  DRB_REQUIRE(e->committed, "drb_engine_encode: parameters not committed");
  DRB_REQUIRE(io->src_grid && io->tgt_grid && io->src_mask && io->tgt_mask, "drb_engine_encode: null tensor");
  DRB_REQUIRE(io->n_src_mask > 0 && io->n_tgt_mask > 0, "drb_engine_encode: empty mask");
  DRB_REQUIRE(io->n_src_mask <= e->cfg.max_mask && io->n_tgt_mask <= e->cfg.max_mask,
              "drb_engine_encode: mask larger than max_mask=%d", e->cfg.max_mask);
  const Act& p1 = e->p[0];
  const int X = e->cfg.res_x, Y = e->cfg.res_y, Z = e->cfg.res_z;
  if (e->grad_mode) {
    DRB_TRY(engine_ensure_grad_buffers(e));
    DRB_REQUIRE(e->committed, "drb_engine_encode: parameters not committed after grad mode was switched on");
  }
  e->graph_valid = false;
  if (e->profile) e->prof_forwards += 1;
  if (e->sparse_fpn) {
    const long long* masks[2] = {io->src_mask, io->tgt_mask};
    const int ks[2] = {io->n_src_mask, io->n_tgt_mask};
    e->launches += 4;
    DRB_TRY(drb_fpn_need_tiles(masks, ks, kG, X, Y, Z, p1.d, p1.h, p1.w, e->need, e->tiles_out, e->tiles_in,
                               e->tile_counts, e->profile ? e->tile_totals : nullptr, s));
    if (e->grad_mode) {   // the backward pass of pyramid_transformation_1 reaches one voxel further
      e->launches += 1;
      DRB_TRY(drb_fpn_dilated_tiles(e->need, kG, p1.d, p1.h, p1.w, 2, e->tiles_in2, e->tile_counts + 2,
                                    e->profile ? e->tile_totals + 2 : nullptr, s));
    }
  }
  DRB_TRY(run_fpn(e, io, s));
  e->launches += 2;
  DRB_TRY(drb_trilinear_gather(p1.f, p1.d, p1.h, p1.w, 256, io->src_grid, io->s_ch, io->s_z, io->s_x,
                               io->s_y, X, Y, Z, io->src_mask, io->n_src_mask, e->rows, kRowLd, s));
  DRB_TRY(drb_trilinear_gather(p1.f + p1.m() * 256, p1.d, p1.h, p1.w, 256, io->tgt_grid, io->t_ch, io->t_z,
                               io->t_x, io->t_y, X, Y, Z, io->tgt_mask, io->n_tgt_mask,
                               e->rows + (long long)io->n_src_mask * kRowLd, kRowLd, s));
  // dl0 = 2 * (0.025 * 2.75) / 2.75 computed in double like grid_downsample.py:68,77
  const double dl0 = 2.0 * (0.025 * 2.75) / 2.75;
  e->launches += 6 * e->cfg.num_downsample;
  e->n_src_mask = io->n_src_mask; e->n_tgt_mask = io->n_tgt_mask;
  e->ds_tape.clear();
  if (e->grad_mode) {
    int info[64], rounds = 0;
    DRB_REQUIRE(e->cfg.num_downsample <= 32, "drb_engine_encode: num_downsample > 32");
    DRB_TRY(drb_hierarchical_downsample_tape(e->rows, io->n_src_mask, io->n_tgt_mask, kRowLd, e->cfg.num_downsample,
                                             dl0, e->max_tokens, e->ds_ws, e->ds_ws_bytes, e->rows_ds, &e->n_src,
                                             &e->n_tgt, e->ds_tape_buf, e->ds_tape_cap, info, &rounds, s));
    long long used = 0;
    for (int r = 0; r < rounds; ++r) {
      DsRound d;
      d.n_in = info[2 * r]; d.n_seg = info[2 * r + 1];
      d.sorted_rows = e->ds_tape_buf + used;
      d.seg_start = d.sorted_rows + d.n_in;
      used += (long long)d.n_in + d.n_seg + 1;
      e->ds_tape.push_back(d);
    }
  } else {
    DRB_TRY(drb_hierarchical_downsample(e->rows, io->n_src_mask, io->n_tgt_mask, kRowLd, e->cfg.num_downsample,
                                        dl0, e->max_tokens, e->ds_ws, e->ds_ws_bytes, e->rows_ds, &e->n_src,
                                        &e->n_tgt, s));
  }
  // the down-sampler synchronised THIS stream: the mask range check of the gather kernels is readable for free
  // (no device-wide wait: other streams may be registering other pairs, pipeline.py)
  const int flag = igemm_peek_err_flag();
  if (flag == 21) {
    igemm_clear_err_flag();
    set_error("drb_engine_encode: a mask index is outside [0, X*Y*Z) (mask of another resolution?)");
    return DRB_EINVAL;
  }
  *host_n_src = e->n_src;
  *host_n_tgt = e->n_tgt;
  return 0;
}

extern "C" int drb_engine_decode(drb_engine* e, const drb_pair_out* o, cudaStream_t s) {
  DRB_REQUIRE(e && o, "drb_engine_decode: null argument");
  DRB_REQUIRE(o->src_feats && o->tgt_feats && o->src_kp && o->tgt_kp && o->src_corr && o->tgt_corr &&
                  o->src_overlap && o->tgt_overlap && o->pose,
              "drb_engine_decode: null output");
  const int ns = e->n_src, nt = e->n_tgt, m = ns + nt;
  DRB_REQUIRE(ns > 0 && nt > 0, "drb_engine_decode: encode() produced no tokens (%d, %d)", ns, nt);
  DRB_TRY(engine_ensure_tokens(e, m));
  e->launches += 2;
  // in grad mode every layer keeps its own copies (TSave / xs); otherwise they all alias one scratch set
  unpack_rows_kernel<<<cdiv((long long)m * 64, 256), 256, 0, s>>>(e->rows_ds, m, e->kp_xyz, e->xs[0]);
  DRB_LAUNCH_OK();
  DRB_TRY(drb_pos_embed_sine(e->kp_xyz, 3, m, e->cfg.pos_emb_scaling, e->pos, s));
  const float att_scale = 1.f / sqrtf(32.f);
  auto linear = [&](const ConvW& w, const plane_t* in_hi, const plane_t* in_lo, int rows, int cin, const float* res,
                    int relu, float scale, float* out, plane_t* ohi, plane_t* olo) {
    return run_igemm(e, w, in_hi, in_lo, 1, 1, 1, rows, cin, 1, P(e, w.p_b), res, relu, scale, out, ohi, olo, 0, s);
  };
  // multi-head attention of the query segment [q0, q0 + nq) against the key segment [k0, k0 + nk) of one in_proj
  // output; `packed` says whether this qkv has already been packed for the tensor-core kernel
  // multi-head attention of both clouds over one in_proj output: self (cross == 0) or against the other cloud
  auto attend = [&](const float* qkv, int cross, plane_t* ohi, plane_t* olo) -> int {
    if (e->tc_attention) {
      e->launches += 2;
      DRB_TRY(drb_mha_tc_pack(qkv, 768, qkv + 256, 768, qkv + 512, 768, m, ns, 8, e->cfg.planes, att_scale, e->att_ws,
                              e->att_ws_bytes, s));
      return drb_mha_tc_forward(e->att_ws, m, ns, 8, e->cfg.planes, -1, cross, nullptr, ohi, olo, 256, s);
    }
    for (int q_seg = 0; q_seg < 2; ++q_seg) {
      const int k_seg = cross ? 1 - q_seg : q_seg;
      const int q0 = q_seg ? ns : 0, nq = q_seg ? nt : ns, k0 = k_seg ? ns : 0, nk = k_seg ? nt : ns;
      e->launches += 1;
      DRB_TRY(drb_mha_core(qkv + (long long)q0 * 768, 768, qkv + (long long)k0 * 768 + 256, 768,
                           qkv + (long long)k0 * 768 + 512, 768, nq, nk, 8, att_scale, nullptr,
                           off(ohi, (long long)q0 * 256), off(olo, (long long)q0 * 256), 256, s));
    }
    return 0;
  };
  for (int l = 0; l < kLayers; ++l) {
    TLayer& t = e->tl[l];
    TSave& v = e->ts[l];
    float* x0 = e->xs[l];
    float* x3 = e->xs[l + 1];
    // self attention (shared weights for src and tgt)
    e->launches += 1;
    DRB_TRY(drb_layernorm256(x0, m, P(e, t.n1w), P(e, t.n1b), e->pos, nullptr, v.xn1_hi, v.xn1_lo, s));
    DRB_TRY(linear(t.self_attn.in_proj, v.xn1_hi, v.xn1_lo, m, 256, nullptr, 0, 1.f, v.qkv_s, nullptr, nullptr));
    DRB_TRY(attend(v.qkv_s, 0, v.att_s_hi, v.att_s_lo));
    DRB_TRY(linear(t.self_attn.out_proj, v.att_s_hi, v.att_s_lo, m, 256, x0, 0, 1.f, v.x1, nullptr, nullptr));
    // cross attention, both directions from the same pre-update normalised features
    e->launches += 1;
    DRB_TRY(drb_layernorm256(v.x1, m, P(e, t.n2w), P(e, t.n2b), e->pos, nullptr, v.xn2_hi, v.xn2_lo, s));
    DRB_TRY(linear(t.cross_attn.in_proj, v.xn2_hi, v.xn2_lo, m, 256, nullptr, 0, 1.f, v.qkv_c, nullptr, nullptr));
    DRB_TRY(attend(v.qkv_c, 1, v.att_c_hi, v.att_c_lo));
    DRB_TRY(linear(t.cross_attn.out_proj, v.att_c_hi, v.att_c_lo, m, 256, v.x1, 0, 1.f, v.x2, nullptr, nullptr));
    // feed forward
    e->launches += 1;
    DRB_TRY(drb_layernorm256(v.x2, m, P(e, t.n3w), P(e, t.n3b), nullptr, nullptr, v.xn3_hi, v.xn3_lo, s));
    DRB_TRY(linear(t.lin1, v.xn3_hi, v.xn3_lo, m, 256, nullptr, 1, 1.f, nullptr, v.ffn_hi, v.ffn_lo));
    DRB_TRY(linear(t.lin2, v.ffn_hi, v.ffn_lo, m, 1024, v.x2, 0, 1.f, x3, nullptr, nullptr));
    // shared final norm -> per-layer outputs (transformer.py:69-84)
    float* sf = o->src_feats + (long long)l * ns * kD;
    float* tf = o->tgt_feats + (long long)l * nt * kD;
    e->launches += 5;
    DRB_TRY(drb_layernorm256(x3, ns, P(e, e->fin_w), P(e, e->fin_b), nullptr, sf, nullptr, nullptr, s));
    DRB_TRY(drb_layernorm256(x3 + (long long)ns * kD, nt, P(e, e->fin_w), P(e, e->fin_b), nullptr, tf, nullptr,
                             nullptr, s));
    DRB_TRY(drb_layernorm256(x3, m, P(e, e->fin_w), P(e, e->fin_b), e->pos, nullptr,
                             e->dec_hi + (long long)l * m * kD, off(e->dec_lo, (long long)l * m * kD), s));
    DRB_TRY(drb_overlap_sigmoid(sf, ns, P(e, e->conf_w), P(e, e->conf_b), o->src_overlap + (long long)l * ns, s));
    DRB_TRY(drb_overlap_sigmoid(tf, nt, P(e, e->conf_w), P(e, e->conf_b), o->tgt_overlap + (long long)l * nt, s));
  }
  // decoder: q / k projections for all layers at once, then per layer soft correspondences
  const int m6 = kLayers * m;
  DRB_TRY(linear(e->q_proj, e->dec_hi, e->dec_lo, m6, 256, nullptr, 0, 1.f / sqrtf((float)kD), e->qf, e->qp_hi,
                 e->qp_lo));
  DRB_TRY(linear(e->k_proj, e->dec_hi, e->dec_lo, m6, 256, nullptr, 0, 1.f, e->kf, e->kp_hi, e->kp_lo));
  e->launches += 1;
  DRB_CUDA_OK(cudaMemcpyAsync(o->src_kp, e->kp_xyz, (size_t)ns * 3 * 4, cudaMemcpyDeviceToDevice, s));
  DRB_CUDA_OK(cudaMemcpyAsync(o->tgt_kp, e->kp_xyz + (long long)ns * 3, (size_t)nt * 3 * 4,
                              cudaMemcpyDeviceToDevice, s));
  const long long ld_t = ((nt + 7) / 8) * 8, ld_s = ((ns + 7) / 8) * 8;
  for (int l = 0; l < kLayers; ++l) {
    const long long base = (long long)l * m * kD;
    // src queries against tgt keys
    DRB_TRY(run_igemm(e, e->q_proj, e->qp_hi + base, off(e->qp_lo, base), 1, 1, 1, ns, 256, 1, nullptr, nullptr, 0,
                      1.f, e->sbuf, nullptr, nullptr, ld_t, s, e->kp_hi + base + (long long)ns * kD,
                      off(e->kp_lo, base + (long long)ns * kD), nt));
    e->launches += 1;
    DRB_TRY(drb_softmax_weighted_xyz(e->sbuf, (int)ld_t, ns, nt, o->tgt_kp, 3,
                                     o->src_corr + (long long)l * ns * 3, s));
    // tgt queries against src keys
    DRB_TRY(run_igemm(e, e->q_proj, e->qp_hi + base + (long long)ns * kD, off(e->qp_lo, base + (long long)ns * kD), 1,
                      1, 1, nt, 256, 1, nullptr, nullptr, 0, 1.f, e->sbuf, nullptr, nullptr, ld_s, s,
                      e->kp_hi + base, off(e->kp_lo, base), ns));
    e->launches += 1;
    DRB_TRY(drb_softmax_weighted_xyz(e->sbuf, (int)ld_s, nt, ns, o->src_kp, 3,
                                     o->tgt_corr + (long long)l * nt * 3, s));
  }
  e->launches += 1;
  DRB_TRY(drb_procrustes(o->src_kp, 0, o->src_corr, (long long)ns * 3, o->src_overlap, ns, ns, o->tgt_corr,
                         (long long)nt * 3, o->tgt_kp, 0, o->tgt_overlap, nt, nt, 3, kLayers, o->pose, s));
  e->graph_valid = e->grad_mode;
  return 0;
}

extern "C" int drb_engine_tap(drb_engine* e, const char* name, int which, float* dst, long long capacity,
                              long long* numel, cudaStream_t s) {
  DRB_REQUIRE(e && name && dst && numel, "drb_engine_tap: null argument");
  const float* src = nullptr;
  long long n = 0;
  const std::string nm(name);
  auto act = [&](const Act& a) { n = a.m() * a.c; src = a.f ? a.f + (long long)which * n : nullptr; };
  if (nm == "c1") act(e->c1);
  else if (nm == "x0") act(e->x0);
  else if (nm.size() == 2 && nm[0] == 'c' && nm[1] >= '2' && nm[1] <= '5') act(e->c[nm[1] - '2']);
  else if (nm.size() == 2 && nm[0] == 'p' && nm[1] >= '1' && nm[1] <= '5') act(e->p[nm[1] - '1']);
  else if (nm == "rows") { src = e->rows; n = capacity; }
  else if (nm == "rows_ds") { src = e->rows_ds; n = (long long)(e->n_src + e->n_tgt) * kRowLd; }
  DRB_REQUIRE(src != nullptr, "drb_engine_tap: unknown or unavailable tap '%s'", name);
  DRB_REQUIRE(n <= capacity, "drb_engine_tap: capacity %lld < %lld", capacity, n);
  DRB_CUDA_OK(cudaMemcpyAsync(dst, src, (size_t)n * sizeof(float), cudaMemcpyDeviceToDevice, s));
  *numel = n;
  return 0;
}
