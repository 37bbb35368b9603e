// Whole-path engine: NeRFRegTr.forward (conerf/register/nerf_regtr.py:112-248) for one (src, tgt)
// pair without Python in the loop.  The two grids of a pair run through the FPN as g = 2 images of
// one launch sequence (BatchNorm statistics stay per grid, as in the reference's two B = 1 calls).
//
// Network topology restated from conerf/model/resnet3d.py:116-172 (ResNet-50, Bottleneck
// [3,4,6,3], conv1 5^3 s2, maxpool 3 s2) and conerf/model/feature_pyramid_net.py:39-108 (FPN v1),
// conerf/register/transformer.py:225-299 (pre-LN cross encoder), nerf_regtr.py:350-394 (decoder).
#include "common.cuh"

#include <string>
#include <vector>

namespace drb {

static constexpr int kG = 2;            // src + tgt
static constexpr int kRowLd = 260;      // [x y z 0 | 256 features]
static constexpr int kLayers = 6;
static constexpr int kD = 256;

struct Act {                 // channels-last activation [g][d][h][w][c]
  float* f = nullptr;
  plane_t* hi = nullptr;
  plane_t* lo = nullptr;
  int d = 0, h = 0, w = 0, c = 0;
  long long m() const { return (long long)d * h * w; }          // voxels per grid
  long long numel() const { return (long long)kG * m() * c; }
};

enum ParamKind { PK_CONV, PK_VEC };

struct Param {
  std::string name;
  long long numel = 0;
  float* ptr = nullptr;
};

struct ConvW {               // packed weight planes
  int p_w = -1, p_b = -1;    // param indices (bias optional)
  int cout = 0, cin = 0, k = 1, stride = 1;
  float scale = 1.f;         // power-of-two pre-scale applied when packing (pair mode)
  bool im2col = false;       // lowered through an explicit im2col buffer
  int kpad = 0;              // im2col K (multiple of 64)
  plane_t* hi = nullptr;
  plane_t* lo = nullptr;
};

struct BnP {
  int p_w = -1, p_b = -1, p_rm = -1, p_rv = -1;
  int c = 0;
  float* scale = nullptr;    // [g][c]
  float* shift = nullptr;
};

struct Block {
  ConvW conv1, conv2, conv3, down;
  BnP bn1, bn2, bn3, bnd;
  bool has_down = false;
  int stride = 1;
};

struct AttnW { ConvW in_proj, out_proj; };
struct TLayer {
  AttnW self_attn, cross_attn;
  ConvW lin1, lin2;
  int n1w, n1b, n2w, n2b, n3w, n3b;
};

}  // namespace drb

using namespace drb;

struct drb_engine {
  drb_engine_config cfg;
  std::vector<Param> params;
  std::vector<void*> allocs;
  long long launches = 0;
  bool committed = false;
  bool profile = false;
  struct ProfRec { cudaEvent_t a, b; double flops; int list_id; double flops_per_tile; };
  std::vector<ProfRec> prof;
  std::string fail;

  // topology
  ConvW conv1; BnP bn1;
  std::vector<Block> blocks[4];
  ConvW pyr[5], ups[4];
  TLayer tl[kLayers];
  int fin_w, fin_b;                 // transformer_encoder.norm
  ConvW q_proj, k_proj;
  int conf_w, conf_b;

  // FPN buffers
  int D, H, W;                      // conv volume axes: D = Z, H = X, W = Y
  plane_t *col_hi = nullptr, *col_lo = nullptr;   // shared im2col scratch
  long long col_elems = 0;
  float* raw = nullptr;             // shared raw conv output scratch (largest BN'd conv)
  long long raw_elems = 0;
  float* raw2 = nullptr;            // second scratch (downsample branch)
  double* bn_accum = nullptr;
  Act c1, x0, c[4];                 // c[0..3] = c2..c5
  std::vector<Act> tmp;             // per-block temporaries
  Act lat[5], sum[4], p[5];         // p[0] = p1 ... p[4] = p5
  // output-sparse evaluation of the two level-1 FPN convolutions
  bool sparse_fpn = true;
  uint8_t* need = nullptr;
  int *tiles_out = nullptr, *tiles_in = nullptr, *tile_counts = nullptr;
  unsigned long long* tile_totals = nullptr;   // running sums of the list lengths (profiling)
  // point stage
  float* rows = nullptr;            // [2*max_mask][260]
  float* rows_ds = nullptr;
  void* ds_ws = nullptr; size_t ds_ws_bytes = 0;
  int n_src = 0, n_tgt = 0;         // tokens after down-sampling
  // transformer buffers (capacity tok_cap rows)
  int tok_cap = 0;
  float *x = nullptr, *pos = nullptr, *qkv = nullptr, *kp_xyz = nullptr, *sbuf = nullptr;
  plane_t *xn_hi = nullptr, *xn_lo = nullptr, *att_hi = nullptr, *att_lo = nullptr;
  plane_t *ffn_hi = nullptr, *ffn_lo = nullptr, *dec_hi = nullptr, *dec_lo = nullptr;
  plane_t *qp_hi = nullptr, *qp_lo = nullptr, *kp_hi = nullptr, *kp_lo = nullptr;
  std::vector<void*> tok_allocs;

  template <typename T> T* alloc(long long n) {
    void* p = nullptr;
    if (n <= 0) n = 1;
    if (cudaMalloc(&p, (size_t)n * sizeof(T)) != cudaSuccess) {
      fail = "cudaMalloc failed";
      return nullptr;
    }
    allocs.push_back(p);
    return (T*)p;
  }
  int add_param(const std::string& name, long long numel) {
    Param q; q.name = name; q.numel = numel;
    params.push_back(q);
    return (int)params.size() - 1;
  }
};

namespace drb {

static ConvW make_conv(drb_engine* e, const std::string& name, int cout, int cin, int k, int stride,
                       bool bias) {
  ConvW w;
  w.cout = cout; w.cin = cin; w.k = k; w.stride = stride;
  w.p_w = e->add_param(name + ".weight", (long long)cout * cin * k * k * k);
  if (bias) w.p_b = e->add_param(name + ".bias", cout);
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
  b.p_w = e->add_param(name + ".weight", c);
  b.p_b = e->add_param(name + ".bias", c);
  b.p_rm = e->add_param(name + ".running_mean", c);
  b.p_rv = e->add_param(name + ".running_var", c);
  b.scale = e->alloc<float>((long long)kG * c);
  b.shift = e->alloc<float>((long long)kG * c);
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
      a.in_proj.cout = 768; a.in_proj.cin = 256;
      a.in_proj.p_w = e->add_param(pn + "." + an + ".in_proj_weight", 768 * 256);
      a.in_proj.p_b = e->add_param(pn + "." + an + ".in_proj_bias", 768);
      a.in_proj.hi = e->alloc<plane_t>(768 * 256);
      if (e->cfg.planes == 2) a.in_proj.lo = e->alloc<plane_t>(768 * 256);
      a.out_proj = make_conv(e, pn + "." + an + ".out_proj", 256, 256, 1, 1, true);
      return a;
    };
    t.self_attn = attn("self_attn");
    t.cross_attn = attn("cross_attn");
    t.lin1 = make_conv(e, pn + ".linear1", 1024, 256, 1, 1, true);
    t.lin2 = make_conv(e, pn + ".linear2", 256, 1024, 1, 1, true);
    t.n1w = e->add_param(pn + ".norm1.weight", 256); t.n1b = e->add_param(pn + ".norm1.bias", 256);
    t.n2w = e->add_param(pn + ".norm2.weight", 256); t.n2b = e->add_param(pn + ".norm2.bias", 256);
    t.n3w = e->add_param(pn + ".norm3.weight", 256); t.n3b = e->add_param(pn + ".norm3.bias", 256);
  }
  e->fin_w = e->add_param("transformer_encoder.norm.weight", 256);
  e->fin_b = e->add_param("transformer_encoder.norm.bias", 256);
  e->q_proj = make_conv(e, "correspondence_decoder.q_proj", 256, 256, 1, 1, true);
  e->k_proj = make_conv(e, "correspondence_decoder.k_proj", 256, 256, 1, 1, true);
  e->conf_w = e->add_param("correspondence_decoder.conf_logits_decoder.weight", 256);
  e->conf_b = e->add_param("correspondence_decoder.conf_logits_decoder.bias", 1);

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
    e->tile_counts = e->alloc<int>(2);
    e->tile_totals = e->alloc<unsigned long long>(2);
    cudaMemset(e->tile_totals, 0, 2 * sizeof(unsigned long long));
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
  return e->fail.empty() ? 0 : DRB_ENOMEM;
}

// --------------------------------------------------------------------------------------------
#define DRB_TRY(expr)            \
  do {                           \
    int _rc = (expr);            \
    if (_rc != 0) return _rc;    \
  } while (0)

static int run_igemm(drb_engine* e, const ConvW& w, const plane_t* x_hi, const plane_t* x_lo, int g, int d,
                     int h, int wd, int cin, int k, const float* bias, const float* residual,
                     int relu, float scale, float* out, plane_t* out_hi, plane_t* out_lo, long long ld,
                     cudaStream_t s, const plane_t* w_hi = nullptr, const plane_t* w_lo = nullptr,
                     int cout_override = 0, const int* tile_list = nullptr, const int* tile_count = nullptr,
                     double* bn_accum = nullptr) {
  drb_conv3d_desc cd;
  memset(&cd, 0, sizeof(cd));
  cd.g = g; cd.d = d; cd.h = h; cd.w = wd;
  cd.cin = cin; cd.cout = cout_override ? cout_override : w.cout;
  cd.kd = cd.kh = cd.kw = k;
  cd.planes = e->cfg.planes;
  cd.relu = relu; cd.out_scale = scale;
  cd.acc_scale = w_hi ? 1.f : 1.f / w.scale;
  cd.x_hi = x_hi; cd.x_lo = x_lo;
  cd.w_hi = w_hi ? w_hi : w.hi; cd.w_lo = w_lo ? w_lo : w.lo;
  cd.bias = bias; cd.residual = residual;
  cd.out = out; cd.out_hi = out_hi; cd.out_lo = out_lo;
  cd.ld_out = ld;
  cd.tile_list = tile_list; cd.tile_count = tile_count;
  cd.bn_accum = bn_accum;
  e->launches += 1;
  if (!e->profile) return drb_conv3d_igemm(&cd, s);
  drb_engine::ProfRec r;
  cudaEventCreate(&r.a);
  cudaEventCreate(&r.b);
  r.flops = 2.0 * (double)g * d * h * wd * (double)cd.cout * (double)cin * k * k * k;
  r.list_id = tile_list == nullptr ? -1 : (tile_list == e->tiles_out ? 0 : 1);
  r.flops_per_tile = 2.0 * 128.0 * (double)cd.cout * (double)cin * k * k * k;   // executed work of one listed tile
  cudaEventRecord(r.a, s);
  const int rc = drb_conv3d_igemm(&cd, s);
  cudaEventRecord(r.b, s);
  e->prof.push_back(r);
  return rc;
}

static inline float* P(drb_engine* e, int idx) { return idx >= 0 ? e->params[idx].ptr : nullptr; }
static inline plane_t* off(plane_t* p, long long n) { return p ? p + n : nullptr; }

// conv (stride 1 via TMA implicit GEMM, otherwise im2col + 1x1x1 GEMM) into raw fp32 [g][m][cout]
static int conv_any(drb_engine* e, const ConvW& w, const Act& in, int od, int oh, int ow, float* out,
                    cudaStream_t s) {
  // Every caller feeds a BatchNorm.  The GEMM epilogue CAN produce the per-channel sums
  // (drb_conv3d_desc.bn_accum, parity-tested), but its fp64 atomics were measured slower than the separate
  // HBM-bound pass on B200 (+1.2 ms vs -0.85 ms per pair, round 1), so the engine keeps the separate pass.
  double* acc = nullptr;
  if (!w.im2col) {
    return run_igemm(e, w, in.hi, in.lo, kG, in.d, in.h, in.w, w.cin, w.k, P(e, w.p_b), nullptr, 0,
                     1.f, out, nullptr, nullptr, 0, s, nullptr, nullptr, 0, nullptr, nullptr, acc);
  }
  drb_im2col_desc d;
  memset(&d, 0, sizeof(d));
  d.x = in.f;
  d.sc = 1; d.sw = in.c; d.sh = (long long)in.w * in.c; d.sd = (long long)in.h * in.w * in.c;
  d.sg = (long long)in.d * in.h * in.w * in.c;
  d.g = kG; d.c = in.c; d.d = in.d; d.h = in.h; d.w = in.w;
  d.k = w.k; d.stride = w.stride; d.pad = w.k / 2; d.kpad = w.kpad;
  e->launches += 1;
  DRB_TRY(drb_im2col(&d, e->col_hi, e->col_lo, s));
  return run_igemm(e, w, e->col_hi, e->col_lo, kG, od, oh, ow, w.kpad, 1, P(e, w.p_b), nullptr, 0, 1.f,
                   out, nullptr, nullptr, 0, s, nullptr, nullptr, 0, nullptr, nullptr, acc);
}

// BatchNorm (+ residual, ReLU) of raw [g][m][c] into out (fp32 and/or planes)
static int bn_apply(drb_engine* e, const BnP& b, const float* rawp, long long m, const float* residual,
                    int relu, float* out, plane_t* out_hi, plane_t* out_lo, cudaStream_t s) {
  const int training = e->cfg.training_bn;
  if (training) {
    e->launches += 2;
    DRB_TRY(drb_bn_stats(rawp, kG, m, b.c, e->bn_accum, s));
  }
  e->launches += 1;
  return drb_bn_apply(rawp, e->bn_accum, kG, m, b.c, P(e, b.p_w), P(e, b.p_b), P(e, b.p_rm), P(e, b.p_rv), training,
                      0.1f, 1e-5f, residual, relu, out, out_hi, out_lo, s);
}

static int run_fpn(drb_engine* e, const drb_pair_io* io, cudaStream_t s) {
  // ---- conv1 (5^3 s2, Cin 4 = rgba channels 3..6 of the [1,7,Z,X,Y] grid) through im2col ----
  {
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
    DRB_TRY(run_igemm(e, w, e->col_hi, e->col_lo, kG, e->c1.d, e->c1.h, e->c1.w, w.kpad, 1, nullptr,
                      nullptr, 0, 1.f, e->raw, nullptr, nullptr, 0, s));
    DRB_TRY(bn_apply(e, e->bn1, e->raw, e->c1.m(), nullptr, 1, e->c1.f, e->c1.hi, e->c1.lo, s));
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
      DRB_TRY(conv_any(e, b.conv1, *x, x->d, x->h, x->w, e->raw, s));
      DRB_TRY(bn_apply(e, b.bn1, e->raw, t1.m(), nullptr, 1, t1.f, t1.hi, t1.lo, s));
      DRB_TRY(conv_any(e, b.conv2, t1, t2.d, t2.h, t2.w, e->raw, s));
      DRB_TRY(bn_apply(e, b.bn2, e->raw, t2.m(), nullptr, 1, nullptr, t2.hi, t2.lo, s));
      const float* res = x->f;
      if (b.has_down) {
        DRB_TRY(conv_any(e, b.down, *x, out.d, out.h, out.w, e->raw2, s));
        DRB_TRY(bn_apply(e, b.bnd, e->raw2, out.m(), nullptr, 0, e->raw2, nullptr, nullptr, s));
        res = e->raw2;
      }
      DRB_TRY(conv_any(e, b.conv3, t2, out.d, out.h, out.w, e->raw, s));
      DRB_TRY(bn_apply(e, b.bn3, e->raw, out.m(), res, 1, out.f, out.hi, out.lo, s));
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
    DRB_TRY(run_igemm(e, e->pyr[i], f.hi, f.lo, kG, f.d, f.h, f.w, f.c, e->pyr[i].k, P(e, e->pyr[i].p_b),
                      nullptr, 0, 1.f, e->lat[i].f, nullptr, nullptr, 0, s, nullptr, nullptr, 0,
                      sparse ? e->tiles_in : nullptr, sparse ? e->tile_counts + 1 : nullptr));
    const Act& top = e->p[i + 1];
    e->launches += 1;
    DRB_TRY(drb_upsample2_add(top.f, top.d, top.h, top.w, e->lat[i].f, kG, f.d, f.h, f.w, 256, nullptr,
                              e->sum[i].hi, e->sum[i].lo, s));
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

static int ensure_tokens(drb_engine* e, int m) {
  if (m <= e->tok_cap) return 0;
  for (void* p : e->tok_allocs) cudaFree(p);
  e->tok_allocs.clear();
  const long long cap = ((m + 127) / 128) * 128 + 128;
  auto A = [&](long long bytes) -> void* {
    void* p = nullptr;
    if (cudaMalloc(&p, (size_t)bytes) != cudaSuccess) return nullptr;
    e->tok_allocs.push_back(p);
    return p;
  };
  const long long cap_ld = ((cap + 7) / 8) * 8;
  e->x = (float*)A(cap * kD * 4); e->pos = (float*)A(cap * kD * 4);
  e->qkv = (float*)A(cap * 768 * 4); e->kp_xyz = (float*)A(cap * 3 * 4 + 64);
  e->sbuf = (float*)A(cap * cap_ld * 4);
  const bool pair = e->cfg.planes == 2;
  auto AL = [&](long long bytes) -> plane_t* { return pair ? (plane_t*)A(bytes) : nullptr; };
  e->xn_hi = (plane_t*)A(cap * kD * 2); e->xn_lo = AL(cap * kD * 2);
  e->att_hi = (plane_t*)A(cap * kD * 2); e->att_lo = AL(cap * kD * 2);
  e->ffn_hi = (plane_t*)A(cap * 1024 * 2); e->ffn_lo = AL(cap * 1024 * 2);
  e->dec_hi = (plane_t*)A(kLayers * cap * kD * 2); e->dec_lo = AL(kLayers * cap * kD * 2);
  e->qp_hi = (plane_t*)A(kLayers * cap * kD * 2); e->qp_lo = AL(kLayers * cap * kD * 2);
  e->kp_hi = (plane_t*)A(kLayers * cap * kD * 2); e->kp_lo = AL(kLayers * cap * kD * 2);
  const bool lo_ok = !pair || (e->xn_lo && e->att_lo && e->ffn_lo && e->dec_lo && e->qp_lo && e->kp_lo);
  if (!e->x || !e->pos || !e->qkv || !e->kp_xyz || !e->sbuf || !e->xn_hi || !e->att_hi || !e->ffn_hi ||
      !e->dec_hi || !e->qp_hi || !e->kp_hi || !lo_ok) {
    set_error("drb_engine: out of device memory for %d tokens", m);
    e->tok_cap = 0;
    return DRB_ENOMEM;
  }
  e->tok_cap = (int)cap;
  return 0;
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

extern "C" int drb_engine_set_profile(drb_engine* e, int on) {
  DRB_REQUIRE(e, "drb_engine_set_profile: null engine");
  e->profile = on != 0;
  return 0;
}

extern "C" int drb_engine_profile_read(drb_engine* e, double* igemm_ms, double* igemm_flops, long long* n) {
  DRB_REQUIRE(e && igemm_ms && igemm_flops && n, "drb_engine_profile_read: null argument");
  double ms = 0.0, fl = 0.0;
  // output-sparse launches: executed FLOPs = (sum of their tile-list lengths) x FLOPs of one tile
  unsigned long long totals[2] = {0, 0};
  double per_tile[2] = {0.0, 0.0};
  DRB_CUDA_OK(cudaDeviceSynchronize());
  if (e->tile_totals) {
    DRB_CUDA_OK(cudaMemcpy(totals, e->tile_totals, sizeof(totals), cudaMemcpyDeviceToHost));
    DRB_CUDA_OK(cudaMemset(e->tile_totals, 0, sizeof(totals)));
  }
  for (auto& r : e->prof) {
    DRB_CUDA_OK(cudaEventSynchronize(r.b));
    float t = 0.f;
    DRB_CUDA_OK(cudaEventElapsedTime(&t, r.a, r.b));
    ms += t;
    if (r.list_id >= 0) per_tile[r.list_id] = r.flops_per_tile; else fl += r.flops;
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  fl += (double)totals[0] * per_tile[0] + (double)totals[1] * per_tile[1];
  *igemm_ms = ms; *igemm_flops = fl; *n = (long long)e->prof.size();
  e->prof.clear();
  return 0;
}

extern "C" int drb_engine_commit_params(drb_engine* e, cudaStream_t s) {
  DRB_REQUIRE(e, "drb_engine_commit_params: null engine");
  for (const Param& p : e->params)
    DRB_REQUIRE(p.ptr != nullptr, "drb_engine_commit_params: parameter %s is not bound", p.name.c_str());
  const bool pair = e->cfg.planes == 2;
  auto pack = [&](ConvW& w) -> int {
    const int taps = w.k * w.k * w.k;
    w.scale = 1.f;
    if (pair) DRB_TRY(drb_weight_scale(P(e, w.p_w), (long long)w.cout * w.cin * taps, &w.scale, s));
    plane_t* lo = pair ? w.lo : nullptr;
    if (w.im2col)
      return drb_pack_conv_weight_im2col(P(e, w.p_w), w.cout, w.cin, taps, w.kpad, w.scale, w.hi, lo, s);
    return drb_pack_conv_weight(P(e, w.p_w), w.cout, w.cin, taps, w.cin, w.scale, w.hi, lo, s);
  };
  DRB_TRY(pack(e->conv1));
  for (int li = 0; li < 4; ++li)
    for (Block& b : e->blocks[li]) {
      DRB_TRY(pack(b.conv1)); DRB_TRY(pack(b.conv2)); DRB_TRY(pack(b.conv3));
      if (b.has_down) DRB_TRY(pack(b.down));
    }
  for (int i = 0; i < 5; ++i) DRB_TRY(pack(e->pyr[i]));
  for (int i = 0; i < 4; ++i) DRB_TRY(pack(e->ups[i]));
  for (int l = 0; l < kLayers; ++l) {
    TLayer& t = e->tl[l];
    DRB_TRY(pack(t.self_attn.in_proj)); DRB_TRY(pack(t.self_attn.out_proj));
    DRB_TRY(pack(t.cross_attn.in_proj)); DRB_TRY(pack(t.cross_attn.out_proj));
    DRB_TRY(pack(t.lin1)); DRB_TRY(pack(t.lin2));
  }
  DRB_TRY(pack(e->q_proj)); DRB_TRY(pack(e->k_proj));
  e->committed = true;
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
  if (e->sparse_fpn) {
    const long long* masks[2] = {io->src_mask, io->tgt_mask};
    const int ks[2] = {io->n_src_mask, io->n_tgt_mask};
    e->launches += 4;
    DRB_TRY(drb_fpn_need_tiles(masks, ks, kG, X, Y, Z, p1.d, p1.h, p1.w, e->need, e->tiles_out, e->tiles_in,
                               e->tile_counts, e->profile ? e->tile_totals : nullptr, s));
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
  DRB_TRY(drb_hierarchical_downsample(e->rows, io->n_src_mask, io->n_tgt_mask, kRowLd, e->cfg.num_downsample,
                                      dl0, 3000, e->ds_ws, e->ds_ws_bytes, e->rows_ds, &e->n_src, &e->n_tgt, s));
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
  DRB_TRY(ensure_tokens(e, m));
  e->launches += 2;
  unpack_rows_kernel<<<cdiv((long long)m * 64, 256), 256, 0, s>>>(e->rows_ds, m, e->kp_xyz, e->x);
  DRB_LAUNCH_OK();
  DRB_TRY(drb_pos_embed_sine(e->kp_xyz, 3, m, e->cfg.pos_emb_scaling, e->pos, s));
  const float att_scale = 1.f / sqrtf(32.f);
  auto linear = [&](const ConvW& w, const plane_t* in_hi, const plane_t* in_lo, int rows, int cin, const float* res,
                    int relu, float scale, float* out, plane_t* ohi, plane_t* olo) {
    return run_igemm(e, w, in_hi, in_lo, 1, 1, 1, rows, cin, 1, P(e, w.p_b), res, relu, scale, out, ohi, olo, 0, s);
  };
  for (int l = 0; l < kLayers; ++l) {
    TLayer& t = e->tl[l];
    // self attention (shared weights for src and tgt)
    e->launches += 1;
    DRB_TRY(drb_layernorm256(e->x, m, P(e, t.n1w), P(e, t.n1b), e->pos, nullptr, e->xn_hi, e->xn_lo, s));
    DRB_TRY(linear(t.self_attn.in_proj, e->xn_hi, e->xn_lo, m, 256, nullptr, 0, 1.f, e->qkv, nullptr, nullptr));
    e->launches += 2;
    DRB_TRY(drb_mha_core(e->qkv, 768, e->qkv + 256, 768, e->qkv + 512, 768, ns, ns, 8, att_scale, nullptr,
                         e->att_hi, e->att_lo, 256, s));
    DRB_TRY(drb_mha_core(e->qkv + (long long)ns * 768, 768, e->qkv + (long long)ns * 768 + 256, 768,
                         e->qkv + (long long)ns * 768 + 512, 768, nt, nt, 8, att_scale, nullptr,
                         e->att_hi + (long long)ns * 256, off(e->att_lo, (long long)ns * 256), 256, s));
    DRB_TRY(linear(t.self_attn.out_proj, e->att_hi, e->att_lo, m, 256, e->x, 0, 1.f, e->x, nullptr, nullptr));
    // cross attention, both directions from the same pre-update normalised features
    e->launches += 1;
    DRB_TRY(drb_layernorm256(e->x, m, P(e, t.n2w), P(e, t.n2b), e->pos, nullptr, e->xn_hi, e->xn_lo, s));
    DRB_TRY(linear(t.cross_attn.in_proj, e->xn_hi, e->xn_lo, m, 256, nullptr, 0, 1.f, e->qkv, nullptr, nullptr));
    e->launches += 2;
    DRB_TRY(drb_mha_core(e->qkv, 768, e->qkv + (long long)ns * 768 + 256, 768,
                         e->qkv + (long long)ns * 768 + 512, 768, ns, nt, 8, att_scale, nullptr, e->att_hi,
                         e->att_lo, 256, s));
    DRB_TRY(drb_mha_core(e->qkv + (long long)ns * 768, 768, e->qkv + 256, 768, e->qkv + 512, 768, nt, ns, 8,
                         att_scale, nullptr, e->att_hi + (long long)ns * 256, off(e->att_lo, (long long)ns * 256),
                         256, s));
    DRB_TRY(linear(t.cross_attn.out_proj, e->att_hi, e->att_lo, m, 256, e->x, 0, 1.f, e->x, nullptr, nullptr));
    // feed forward
    e->launches += 1;
    DRB_TRY(drb_layernorm256(e->x, m, P(e, t.n3w), P(e, t.n3b), nullptr, nullptr, e->xn_hi, e->xn_lo, s));
    DRB_TRY(linear(t.lin1, e->xn_hi, e->xn_lo, m, 256, nullptr, 1, 1.f, nullptr, e->ffn_hi, e->ffn_lo));
    DRB_TRY(linear(t.lin2, e->ffn_hi, e->ffn_lo, m, 1024, e->x, 0, 1.f, e->x, nullptr, nullptr));
    // shared final norm -> per-layer outputs (transformer.py:69-84)
    float* sf = o->src_feats + (long long)l * ns * kD;
    float* tf = o->tgt_feats + (long long)l * nt * kD;
    e->launches += 5;
    DRB_TRY(drb_layernorm256(e->x, ns, P(e, e->fin_w), P(e, e->fin_b), nullptr, sf, nullptr, nullptr, s));
    DRB_TRY(drb_layernorm256(e->x + (long long)ns * kD, nt, P(e, e->fin_w), P(e, e->fin_b), nullptr, tf, nullptr,
                             nullptr, s));
    DRB_TRY(drb_layernorm256(e->x, m, P(e, e->fin_w), P(e, e->fin_b), e->pos, nullptr,
                             e->dec_hi + (long long)l * m * kD, off(e->dec_lo, (long long)l * m * kD), s));
    DRB_TRY(drb_overlap_sigmoid(sf, ns, P(e, e->conf_w), P(e, e->conf_b), o->src_overlap + (long long)l * ns, s));
    DRB_TRY(drb_overlap_sigmoid(tf, nt, P(e, e->conf_w), P(e, e->conf_b), o->tgt_overlap + (long long)l * nt, s));
  }
  // decoder: q / k projections for all layers at once, then per layer soft correspondences
  const int m6 = kLayers * m;
  DRB_TRY(linear(e->q_proj, e->dec_hi, e->dec_lo, m6, 256, nullptr, 0, 1.f / sqrtf((float)kD), nullptr, e->qp_hi,
                 e->qp_lo));
  DRB_TRY(linear(e->k_proj, e->dec_hi, e->dec_lo, m6, 256, nullptr, 0, 1.f, nullptr, e->kp_hi, e->kp_lo));
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
