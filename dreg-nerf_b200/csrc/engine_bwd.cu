// Backward pass of the whole-path engine: gradients of every trainable parameter of NeRFRegTr given the
// gradients of its outputs (what `loss.backward()` at train_nerf_regtr.py:229 computes through the modules
// of conerf/register/nerf_regtr.py:112-248).  Mirrors engine.cu's forward in reverse; every GEMM-shaped
// gradient runs on the tcgen05 kernels (data gradients: igemm.cu over flipped / transposed weight planes,
// weight gradients: wgrad.cu), everything else is in backward.cu.
#include "engine.cuh"

using namespace drb;

namespace drb {

static inline int conv_out(int n, int k, int s, int p) { return (n + 2 * p - k) / s + 1; }

int engine_ensure_grad_buffers(drb_engine* e) {
  if (e->grad_allocated) return 0;
  const bool pair = e->cfg.planes == 2;
  // data-gradient weight planes
  for (ConvW* w : e->all_convs) {
    if (!w->need_dgrad) continue;
    const int taps = w->k * w->k * w->k;
    const long long n = w->im2col ? (long long)w->cout * w->kpad : (long long)taps * w->cout * w->cin;
    w->thi = e->alloc<plane_t>(n);
    if (pair) w->tlo = e->alloc<plane_t>(n);
  }
  // per-BatchNorm copies of the convolution outputs
  auto keep = [&](BnP& b, long long elems) { b.raw_elems = elems; b.raw_keep = e->alloc<float>(elems); };
  keep(e->bn1, e->c1.numel());
  long long dcol_max = 1;
  size_t ti = 0;
  const Act* x = &e->x0;
  for (int li = 0; li < 4; ++li)
    for (Block& b : e->blocks[li]) {
      const Act& t1 = e->tmp[ti]; const Act& t2 = e->tmp[ti + 1]; const Act& out = e->tmp[ti + 2];
      ti += 3;
      keep(b.bn1, t1.numel()); keep(b.bn2, t2.numel()); keep(b.bn3, out.numel());
      if (b.has_down) keep(b.bnd, out.numel());
      if (b.conv2.im2col) dcol_max = std::max(dcol_max, (long long)kG * t2.m() * b.conv2.kpad);
      if (b.has_down && b.down.im2col) dcol_max = std::max(dcol_max, (long long)kG * out.m() * b.down.kpad);
      x = &out;
    }
  (void)x;
  e->dcol_elems = dcol_max;
  e->dcol = e->alloc<float>(dcol_max);
  // gradient scratch, sized by the largest activation (p1 / the level-1 lateral)
  long long g = e->p[0].numel();
  g = std::max(g, e->c1.numel());
  for (const Act& a : e->tmp) g = std::max(g, a.numel());
  g = std::max(g, (long long)2 * e->cfg.max_mask * kD);
  e->g_elems = g;
  e->gA = e->alloc<float>(g); e->gB = e->alloc<float>(g);
  e->gT1 = e->alloc<float>(g); e->gT2 = e->alloc<float>(g);
  e->gp_hi = e->alloc<plane_t>(g);
  if (pair) e->gp_lo = e->alloc<plane_t>(g);
  const Act* feats[5] = {&e->c1, &e->c[0], &e->c[1], &e->c[2], &e->c[3]};
  for (int i = 0; i < 5; ++i) e->dF[i] = e->alloc<float>(feats[i]->numel());
  e->bn_sums = e->alloc<double>((long long)kG * 2048 * 2);
  // staging buffer of the weight-gradient kernel: the largest [Cout][taps * Cin] (or [Cout][kpad]) result
  long long stage = 1;
  for (const ConvW* w : e->all_convs)
    if (w->im2col || w->k > 1)
      stage = std::max(stage, (long long)w->cout * (w->im2col ? w->kpad : w->k * w->k * w->k * w->cin));
  e->wg_stage_elems = stage;
  e->wg_stage = e->alloc<float>(stage);
  e->ds_tape_cap = (long long)e->cfg.num_downsample * (4LL * e->cfg.max_mask + 2);
  e->ds_tape_buf = e->alloc<int>(e->ds_tape_cap);
  if (!e->fail.empty()) {
    set_error("drb_engine: out of device memory for the training graph");
    return DRB_ENOMEM;
  }
  e->grad_allocated = true;
  e->committed = false;
  return 0;
}

// ring of device scale slots for gradient tensors
static float* next_slot(drb_engine* e) {
  const int nw = (int)e->all_convs.size();
  float* s = e->slots + 2 * (nw + (e->grad_slot_next % 64));
  e->grad_slot_next += 1;
  return s;
}

struct GradPlanes {
  const plane_t* hi;
  const plane_t* lo;
  const float* inv_scale;   // device scalar or null
};

// fp32 gradient [rows][cols] -> planes in the shared plane scratch (or `hi_dst` / `lo_dst`)
static int split_grad(drb_engine* e, const float* x, long long rows, int cols, plane_t* hi_dst, plane_t* lo_dst,
                      GradPlanes& gp, cudaStream_t s) {
  const bool pair = e->cfg.planes == 2;
  float* slot = next_slot(e);
  e->launches += pair ? 2 : 1;
  DRB_TRY(drb_grad_split(x, rows, cols, cols, cols, hi_dst, pair ? lo_dst : nullptr, slot, s));
  gp.hi = hi_dst; gp.lo = pair ? lo_dst : nullptr;
  gp.inv_scale = pair ? slot + 1 : nullptr;
  return 0;
}

static int wgrad(drb_engine* e, const ConvW& w, const GradPlanes& dy, const plane_t* x_hi, const plane_t* x_lo, int g,
                 int d, int h, int wd, int cout, int cin, int k, int c_real, int taps_real, float* dw_base,
                 const int* tile_list, const int* tile_count, float scale, cudaStream_t s,
                 const plane_t* dy_hi_override = nullptr, const plane_t* dy_lo_override = nullptr) {
  (void)w;
  if (!dw_base) return 0;
  drb_wgrad_desc wd_;
  memset(&wd_, 0, sizeof(wd_));
  wd_.g = g; wd_.d = d; wd_.h = h; wd_.w = wd;
  wd_.cout = cout; wd_.cin = cin;
  wd_.kd = wd_.kh = wd_.kw = k;
  wd_.planes = e->cfg.planes;
  wd_.dy_hi = dy_hi_override ? dy_hi_override : dy.hi;
  wd_.dy_lo = dy_hi_override ? dy_lo_override : dy.lo;
  wd_.x_hi = x_hi; wd_.x_lo = x_lo;
  wd_.scale = scale;
  wd_.scale_dev = dy.inv_scale;
  wd_.dw = dw_base;
  wd_.c_real = c_real; wd_.taps_real = taps_real;
  wd_.tile_list = tile_list; wd_.tile_count = tile_count;
  wd_.stage = e->wg_stage; wd_.stage_elems = e->wg_stage_elems;
  e->launches += (k > 1 || c_real != cin) && e->wg_stage ? 2 : 1;      // + the transposition kernel when staged
  if (!e->profile) return drb_conv3d_wgrad(&wd_, s);
  drb_engine::ProfRec r;
  cudaEventCreate(&r.a);
  cudaEventCreate(&r.b);
  r.flops = 2.0 * (double)g * d * h * wd * (double)cout * (double)cin * k * k * k;
  r.list_id = engine_list_id(e, tile_list);
  r.flops_per_tile = 2.0 * 128.0 * (double)cout * (double)cin * k * k * k;
  r.m = -(g * d * h * wd); r.cin = cin; r.cout = cout; r.k = k;      // negative M marks a weight-gradient launch in the dump
  cudaEventRecord(r.a, s);
  const int rc = drb_conv3d_wgrad(&wd_, s);
  cudaEventRecord(r.b, s);
  e->prof.push_back(r);
  return rc;
}

// Backward of one convolution given its output gradient dy (fp32 [g][od*oh*ow][cout]):
//   weight (+ bias) gradient, and - when dx != null - the input gradient (+ residual) into dx.
// `in` is the forward input.  Clobbers the plane scratch, the im2col scratch and dcol.
static int conv_backward(drb_engine* e, const ConvW& w, const float* dy, const Act& in, int od, int oh, int ow,
                         float* dx, const float* dx_residual, const int* tl_w, const int* tc_w, const int* tl_d,
                         const int* tc_d, cudaStream_t s) {
  const long long rows = (long long)kG * od * oh * ow;
  GradPlanes gp;
  DRB_TRY(split_grad(e, dy, rows, w.cout, e->gp_hi, e->gp_lo, gp, s));
  if (w.p_b >= 0 && GRAD(e, w.p_b)) {
    e->launches += 1;
    DRB_TRY(drb_colsum_add(dy, rows, w.cout, w.cout, GRAD(e, w.p_b), s));
  }
  const float* wscale = (e->cfg.planes == 2 && w.slot) ? w.slot + 1 : nullptr;
  const int taps = w.k * w.k * w.k;
  if (!w.im2col) {
    DRB_TRY(wgrad(e, w, gp, in.hi, in.lo, kG, in.d, in.h, in.w, w.cout, w.cin, w.k, w.cin, taps, GRAD(e, w.p_w), tl_w,
                  tc_w, 1.f, s));
    if (dx)
      DRB_TRY(engine_run_igemm(e, w, gp.hi, gp.lo, kG, in.d, in.h, in.w, w.cout, w.k, nullptr, dx_residual, 0, 1.f, dx,
                               nullptr, nullptr, 0, s, w.thi, w.tlo, w.cin, tl_d, tc_d, wscale, gp.inv_scale));
    return 0;
  }
  // strided / narrow convolution: the forward ran a 1x1x1 GEMM over an im2col buffer
  DRB_TRY(engine_im2col(e, w, in, s));
  DRB_TRY(wgrad(e, w, gp, e->col_hi, e->col_lo, kG, od, oh, ow, w.cout, w.kpad, 1, w.cin, taps, GRAD(e, w.p_w), nullptr,
                nullptr, 1.f, s));
  if (dx) {
    DRB_REQUIRE(rows * w.kpad <= e->dcol_elems, "drb_engine_backward: dcol scratch too small");
    DRB_TRY(engine_run_igemm(e, w, gp.hi, gp.lo, kG, od, oh, ow, w.cout, 1, nullptr, nullptr, 0, 1.f, e->dcol, nullptr,
                             nullptr, 0, s, w.thi, w.tlo, w.kpad, nullptr, nullptr, wscale, gp.inv_scale));
    e->launches += 1;
    DRB_TRY(drb_col2im(e->dcol, kG, w.cin, in.d, in.h, in.w, w.k, w.stride, w.k / 2, w.kpad, dx_residual, dx, s));
  }
  return 0;
}

static int bn_backward(drb_engine* e, const BnP& b, float* dy, long long m, const float* post, int relu, float* dx,
                       cudaStream_t s) {
  e->launches += 2;
  return drb_bn_backward(dy, b.raw_keep, post, b.scale, b.shift, b.mean, b.rstd, P(e, b.p_w), relu,
                         e->cfg.training_bn, kG, m, b.c, e->bn_sums, dx, GRAD(e, b.p_w), GRAD(e, b.p_b), s);
}

// Linear layer backward over `rows` tokens: dy fp32 [rows][cout]; x planes [rows][cin].
static int linear_backward(drb_engine* e, const ConvW& w, const float* dy, int rows, const plane_t* x_hi,
                           const plane_t* x_lo, float* dx, const float* dx_residual, cudaStream_t s) {
  GradPlanes gp;
  DRB_TRY(split_grad(e, dy, rows, w.cout, e->t_ph, e->t_pl, gp, s));
  if (w.p_b >= 0 && GRAD(e, w.p_b)) {
    e->launches += 1;
    DRB_TRY(drb_colsum_add(dy, rows, w.cout, w.cout, GRAD(e, w.p_b), s));
  }
  DRB_TRY(wgrad(e, w, gp, x_hi, x_lo, 1, 1, 1, rows, w.cout, w.cin, 1, w.cin, 1, GRAD(e, w.p_w), nullptr, nullptr, 1.f, s));
  if (dx) {
    const float* wscale = (e->cfg.planes == 2 && w.slot) ? w.slot + 1 : nullptr;
    DRB_TRY(engine_run_igemm(e, w, gp.hi, gp.lo, 1, 1, 1, rows, w.cout, 1, nullptr, dx_residual, 0, 1.f, dx, nullptr,
                             nullptr, 0, s, w.thi, w.tlo, w.cin, nullptr, nullptr, wscale, gp.inv_scale));
  }
  return 0;
}

__global__ void pack_feat_grad_kernel(const float* __restrict__ a, int na, const float* __restrict__ b, int nb,
                                      float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)(na + nb) * 64;
  if (i >= total) return;
  const long long row = i >> 6;
  const int c4 = (int)(i & 63) * 4;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (row < na) { if (a) v = *(const float4*)(a + row * 256 + c4); }
  else if (b) v = *(const float4*)(b + (row - na) * 256 + c4);
  *(float4*)(out + row * 256 + c4) = v;
}

__global__ void copy_or_zero_kernel(const float* __restrict__ src, float* __restrict__ dst, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src ? src[i] : 0.f;
}

// ---------------------------------------------------------------------------------------------
// transformer + decoder + Procrustes backward -> gradient of the down-sampled token features in e->t_dx
// ---------------------------------------------------------------------------------------------
static int tokens_backward(drb_engine* e, const drb_pair_out* o, const drb_pair_grad* g, cudaStream_t s) {
  const int ns = e->n_src, nt = e->n_tgt, m = ns + nt;
  const float att_scale = 1.f / sqrtf(32.f);
  // ---- gradients of the soft correspondences and overlaps: caller's + Procrustes' ----
  float* dcorr_s = e->t_dcorr;                              // [6][ns][3]
  float* dcorr_t = e->t_dcorr + (long long)kLayers * ns * 3;  // [6][nt][3]
  float* dov_s = e->t_dov;
  float* dov_t = e->t_dov + (long long)kLayers * ns;
  auto cz = [&](const float* src, float* dst, long long n) {
    copy_or_zero_kernel<<<cdiv(n, 256), 256, 0, s>>>(src, dst, n);
    return cudaGetLastError() == cudaSuccess ? 0 : DRB_ECUDA;
  };
  e->launches += 4;
  DRB_TRY(cz(g->d_src_corr, dcorr_s, (long long)kLayers * ns * 3));
  DRB_TRY(cz(g->d_tgt_corr, dcorr_t, (long long)kLayers * nt * 3));
  DRB_TRY(cz(g->d_src_overlap, dov_s, (long long)kLayers * ns));
  DRB_TRY(cz(g->d_tgt_overlap, dov_t, (long long)kLayers * nt));
  if (g->d_pose) {
    // forward: a = [src_kp ; tgt_corr], b = [src_corr ; tgt_kp], w = [src_overlap ; tgt_overlap]
    e->launches += 1;
    DRB_TRY(drb_procrustes_backward(o->src_kp, 0, o->src_corr, (long long)ns * 3, o->src_overlap, ns, ns, o->tgt_corr,
                                    (long long)nt * 3, o->tgt_kp, 0, o->tgt_overlap, nt, nt, 3, kLayers, g->d_pose,
                                    nullptr, dcorr_s, dov_s, dcorr_t, nullptr, dov_t, s));
  }
  // ---- decoder: corr = softmax(q k^T) xyz_other, q = (Wq dec + bq) / 16, k = Wk dec + bk ----
  const long long ld_t = ((nt + 7) / 8) * 8, ld_s = ((ns + 7) / 8) * 8;
  const float inv_sqrt_d = 1.f / sqrtf((float)kD);
  for (int l = 0; l < kLayers; ++l) {
    const long long base = (long long)l * m * kD;
    float* dq = e->t_dqf + base;
    float* dk = e->t_dkf + base;
    const float* qf = e->qf + base;
    const float* kf = e->kf + base;
    // src queries against tgt keys
    DRB_TRY(engine_run_igemm(e, e->q_proj, e->qp_hi + base, off(e->qp_lo, base), 1, 1, 1, ns, 256, 1, nullptr, nullptr, 0,
                             1.f, e->sbuf, nullptr, nullptr, ld_t, s, e->kp_hi + base + (long long)ns * kD,
                             off(e->kp_lo, base + (long long)ns * kD), nt));
    e->launches += 3;
    DRB_TRY(drb_softmax_weighted_xyz_backward(e->sbuf, (int)ld_t, ns, nt, o->tgt_kp, 3, dcorr_s + (long long)l * ns * 3, s));
    // dq_s = dS k_t / 16 (gradient w.r.t. the un-scaled projection) ; dk_t = dS^T q_s
    DRB_TRY(drb_sgemm_strided(e->sbuf, 0, ld_t, 1, kf + (long long)ns * kD, 0, kD, 1, dq, 0, kD, ns, kD, nt, 1,
                              inv_sqrt_d, 0, s));
    DRB_TRY(drb_sgemm_strided(e->sbuf, 0, 1, ld_t, qf, 0, kD, 1, dk + (long long)ns * kD, 0, kD, nt, kD, ns, 1, 1.f, 0, s));
    // tgt queries against src keys
    DRB_TRY(engine_run_igemm(e, e->q_proj, e->qp_hi + base + (long long)ns * kD, off(e->qp_lo, base + (long long)ns * kD),
                             1, 1, 1, nt, 256, 1, nullptr, nullptr, 0, 1.f, e->sbuf, nullptr, nullptr, ld_s, s,
                             e->kp_hi + base, off(e->kp_lo, base), ns));
    e->launches += 3;
    DRB_TRY(drb_softmax_weighted_xyz_backward(e->sbuf, (int)ld_s, nt, ns, o->src_kp, 3, dcorr_t + (long long)l * nt * 3, s));
    DRB_TRY(drb_sgemm_strided(e->sbuf, 0, ld_s, 1, kf, 0, kD, 1, dq + (long long)ns * kD, 0, kD, nt, kD, ns, 1,
                              inv_sqrt_d, 0, s));
    DRB_TRY(drb_sgemm_strided(e->sbuf, 0, 1, ld_s, qf + (long long)ns * kD, 0, kD, 1, dk, 0, kD, ns, kD, nt, 1, 1.f, 0, s));
  }
  // qf holds the scaled projection: d(q k^T)/dk uses q as stored, d/dq uses k; the 1/16 of q's definition is
  // applied to dq above, so both projections are now plain linear layers over the 6 m decoder inputs.
  const int m6 = kLayers * m;
  DRB_TRY(linear_backward(e, e->q_proj, e->t_dqf, m6, e->dec_hi, e->dec_lo, e->t_ddec, nullptr, s));
  DRB_TRY(linear_backward(e, e->k_proj, e->t_dkf, m6, e->dec_hi, e->dec_lo, e->t_ddec, e->t_ddec, s));

  // ---- encoder layers in reverse; t_dx = gradient of the residual stream ----
  float* dx = e->t_dx;
  DRB_CUDA_OK(cudaMemsetAsync(dx, 0, sizeof(float) * (size_t)m * kD, s));
  for (int l = kLayers - 1; l >= 0; --l) {
    TLayer& t = e->tl[l];
    TSave& v = e->ts[l];
    // shared final norm: feats_l (+ overlap head) and the decoder input
    float* dfin = e->t_dy;
    e->launches += 4;
    pack_feat_grad_kernel<<<cdiv((long long)m * 64, 256), 256, 0, s>>>(
        g->d_src_feats ? g->d_src_feats + (long long)l * ns * kD : nullptr, ns,
        g->d_tgt_feats ? g->d_tgt_feats + (long long)l * nt * kD : nullptr, nt, dfin);
    DRB_LAUNCH_OK();
    DRB_TRY(drb_overlap_sigmoid_backward(o->src_feats + (long long)l * ns * kD, o->src_overlap + (long long)l * ns,
                                         dov_s + (long long)l * ns, ns, P(e, e->conf_w), dfin, GRAD(e, e->conf_w),
                                         GRAD(e, e->conf_b), s));
    DRB_TRY(drb_overlap_sigmoid_backward(o->tgt_feats + (long long)l * nt * kD, o->tgt_overlap + (long long)l * nt,
                                         dov_t + (long long)l * nt, nt, P(e, e->conf_w), dfin + (long long)ns * kD,
                                         GRAD(e, e->conf_w), GRAD(e, e->conf_b), s));
    DRB_TRY(drb_add_inplace(dfin, e->t_ddec + (long long)l * m * kD, (long long)m * kD, s));
    e->launches += 1;
    DRB_TRY(drb_layernorm256_backward(e->xs[l + 1], dfin, m, P(e, e->fin_w), dx, 0, GRAD(e, e->fin_w),
                                      GRAD(e, e->fin_b), s));
    // feed forward: x3 = x2 + W2 relu(W1 LN3(x2) + b1) + b2
    DRB_TRY(linear_backward(e, t.lin2, dx, m, v.ffn_hi, v.ffn_lo, e->t_dh, nullptr, s));
    e->launches += 1;
    DRB_TRY(drb_relu_mask_plane(e->t_dh, v.ffn_hi, (long long)m * 1024, s));
    DRB_TRY(linear_backward(e, t.lin1, e->t_dh, m, v.xn3_hi, v.xn3_lo, e->t_dxn, nullptr, s));
    e->launches += 1;
    DRB_TRY(drb_layernorm256_backward(v.x2, e->t_dxn, m, P(e, t.n3w), dx, 0, GRAD(e, t.n3w), GRAD(e, t.n3b), s));
    // cross attention: x2 = x1 + out_proj(att_c)
    DRB_TRY(linear_backward(e, t.cross_attn.out_proj, dx, m, v.att_c_hi, v.att_c_lo, e->t_datt, nullptr, s));
    {
      const float* q = v.qkv_c;
      float* dq = e->t_dqkv;
      const long long so = (long long)ns * 768, sa = (long long)ns * kD;
      e->launches += 14;
      // src queries, tgt keys / values
      DRB_TRY(drb_mha_core_backward(q, 768, q + so + 256, 768, q + so + 512, 768, e->t_datt, 256, ns, nt, 8, att_scale, dq,
                                    768, dq + so + 256, 768, dq + so + 512, 768, e->mha_ws, e->mha_ws_bytes, s));
      // tgt queries, src keys / values
      DRB_TRY(drb_mha_core_backward(q + so, 768, q + 256, 768, q + 512, 768, e->t_datt + sa, 256, nt, ns, 8, att_scale,
                                    dq + so, 768, dq + 256, 768, dq + 512, 768, e->mha_ws, e->mha_ws_bytes, s));
    }
    DRB_TRY(linear_backward(e, t.cross_attn.in_proj, e->t_dqkv, m, v.xn2_hi, v.xn2_lo, e->t_dxn, nullptr, s));
    e->launches += 1;
    DRB_TRY(drb_layernorm256_backward(v.x1, e->t_dxn, m, P(e, t.n2w), dx, 0, GRAD(e, t.n2w), GRAD(e, t.n2b), s));
    // self attention: x1 = x0 + out_proj(att_s)
    DRB_TRY(linear_backward(e, t.self_attn.out_proj, dx, m, v.att_s_hi, v.att_s_lo, e->t_datt, nullptr, s));
    {
      const float* q = v.qkv_s;
      float* dq = e->t_dqkv;
      const long long so = (long long)ns * 768, sa = (long long)ns * kD;
      e->launches += 14;
      DRB_TRY(drb_mha_core_backward(q, 768, q + 256, 768, q + 512, 768, e->t_datt, 256, ns, ns, 8, att_scale, dq, 768,
                                    dq + 256, 768, dq + 512, 768, e->mha_ws, e->mha_ws_bytes, s));
      DRB_TRY(drb_mha_core_backward(q + so, 768, q + so + 256, 768, q + so + 512, 768, e->t_datt + sa, 256, nt, nt, 8,
                                    att_scale, dq + so, 768, dq + so + 256, 768, dq + so + 512, 768, e->mha_ws,
                                    e->mha_ws_bytes, s));
    }
    DRB_TRY(linear_backward(e, t.self_attn.in_proj, e->t_dqkv, m, v.xn1_hi, v.xn1_lo, e->t_dxn, nullptr, s));
    e->launches += 1;
    DRB_TRY(drb_layernorm256_backward(e->xs[l], e->t_dxn, m, P(e, t.n1w), dx, 0, GRAD(e, t.n1w), GRAD(e, t.n1b), s));
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// down-sampling + gather backward: e->t_dx [m][256] -> dP1 in e->gA (p[0] layout, zero elsewhere)
// ---------------------------------------------------------------------------------------------
static int points_backward(drb_engine* e, const drb_pair_io* io, cudaStream_t s) {
  const float* cur = e->t_dx;
  float* bufs[2] = {e->gT1, e->gT2};
  int which = 0;
  for (int r = (int)e->ds_tape.size() - 1; r >= 0; --r) {
    const DsRound& d = e->ds_tape[r];
    DRB_REQUIRE((long long)d.n_in * kD <= e->g_elems, "drb_engine_backward: gradient scratch too small");
    e->launches += 1;
    DRB_TRY(drb_segment_mean_backward(cur, kD, d.sorted_rows, d.seg_start, d.n_seg, kD, bufs[which], kD, s));
    cur = bufs[which];
    which ^= 1;
  }
  const Act& p1 = e->p[0];
  const int X = e->cfg.res_x, Y = e->cfg.res_y, Z = e->cfg.res_z;
  DRB_CUDA_OK(cudaMemsetAsync(e->gA, 0, sizeof(float) * (size_t)p1.numel(), s));
  e->launches += 3;
  DRB_TRY(drb_trilinear_gather_backward(cur, kD, 0, p1.d, p1.h, p1.w, kD, X, Y, Z, io->src_mask, e->n_src_mask, e->gA, s));
  DRB_TRY(drb_trilinear_gather_backward(cur + (long long)e->n_src_mask * kD, kD, 0, p1.d, p1.h, p1.w, kD, X, Y, Z,
                                        io->tgt_mask, e->n_tgt_mask, e->gA + p1.m() * kD, s));
  return 0;
}

// ---------------------------------------------------------------------------------------------
// FPN + backbone backward: dP1 in e->gA -> all convolution / BatchNorm parameter gradients
// ---------------------------------------------------------------------------------------------
static int fpn_backward(drb_engine* e, const drb_pair_io* io, cudaStream_t s) {
  const Act* feats[5] = {&e->c1, &e->c[0], &e->c[1], &e->c[2], &e->c[3]};
  float* gA = e->gA;      // current output gradient
  float* gB = e->gB;
  // ---- feature pyramid, bottom-up in reverse (feature_pyramid_net.py:63-105) ----
  for (int i = 0; i < 4; ++i) {
    const Act& f = *feats[i];
    const bool sparse = (i == 0) && e->sparse_fpn;
    // p_i = upsample_transform_i(sum_i): dSum_i -> gB.  Sparse level 1: dP1 is non-zero only inside the tiles the
    // gather touched (tiles_out); its data gradient reaches one voxel further (tiles_in).
    Act sumi = e->sum[i];
    if (sparse) DRB_CUDA_OK(cudaMemsetAsync(gB, 0, sizeof(float) * (size_t)sumi.numel(), s));
    DRB_TRY(conv_backward(e, e->ups[i], gA, sumi, f.d, f.h, f.w, gB, nullptr, sparse ? e->tiles_out : nullptr,
                          sparse ? e->tile_counts : nullptr, sparse ? e->tiles_in : nullptr,
                          sparse ? e->tile_counts + 1 : nullptr, s));
    // sum_i = nearest_up(p_{i+1}) + lateral_i: dLateral_i = dSum_i ; dP_{i+1} = 2^3 sum pool of dSum_i -> gA
    if (sparse) DRB_CUDA_OK(cudaMemsetAsync(e->dF[i], 0, sizeof(float) * (size_t)f.numel(), s));
    DRB_TRY(conv_backward(e, e->pyr[i], gB, f, f.d, f.h, f.w, e->dF[i], nullptr, sparse ? e->tiles_in : nullptr,
                          sparse ? e->tile_counts + 1 : nullptr, sparse ? e->tiles_in2 : nullptr,
                          sparse ? e->tile_counts + 2 : nullptr, s));
    const Act& top = e->p[i + 1];
    e->launches += 1;
    DRB_TRY(drb_upsample2_add_backward(gB, kG, f.d, f.h, f.w, 256, top.d, top.h, top.w, gA, s));
  }
  DRB_TRY(conv_backward(e, e->pyr[4], gA, *feats[4], feats[4]->d, feats[4]->h, feats[4]->w, e->dF[4], nullptr, nullptr,
                        nullptr, nullptr, nullptr, s));
  // ---- backbone, last block first (resnet3d.py:76-113,157-172) ----
  DRB_CUDA_OK(cudaMemcpyAsync(gA, e->dF[4], sizeof(float) * (size_t)feats[4]->numel(), cudaMemcpyDeviceToDevice, s));
  float* T1 = e->gT1;
  float* T2 = e->gT2;
  size_t ti = e->tmp.size();
  for (int li = 3; li >= 0; --li) {
    for (int bi = (int)e->blocks[li].size() - 1; bi >= 0; --bi) {
      const Block& b = e->blocks[li][(size_t)bi];
      ti -= 3;
      const Act& t1 = e->tmp[ti]; const Act& t2 = e->tmp[ti + 1]; const Act& out = e->tmp[ti + 2];
      const Act& xin = (ti == 0) ? e->x0 : e->tmp[ti - 1];
      // out = relu(bn3(conv3(t2)) + res): mask in place (gA becomes the residual branch's gradient), dRaw3 -> T1
      DRB_TRY(bn_backward(e, b.bn3, gA, out.m(), out.f, 1, T1, s));
      DRB_TRY(conv_backward(e, b.conv3, T1, t2, out.d, out.h, out.w, T2, nullptr, nullptr, nullptr, nullptr, nullptr, s));
      DRB_TRY(bn_backward(e, b.bn2, T2, t2.m(), nullptr, 1, T2, s));
      DRB_TRY(conv_backward(e, b.conv2, T2, t1, t2.d, t2.h, t2.w, T1, nullptr, nullptr, nullptr, nullptr, nullptr, s));
      DRB_TRY(bn_backward(e, b.bn1, T1, t1.m(), nullptr, 1, T1, s));
      const float* res_grad = gA;
      if (b.has_down) {
        DRB_TRY(bn_backward(e, b.bnd, gA, out.m(), nullptr, 0, T2, s));
        DRB_TRY(conv_backward(e, b.down, T2, xin, out.d, out.h, out.w, gB, nullptr, nullptr, nullptr, nullptr, nullptr, s));
        res_grad = gB;
      }
      // dX = conv1^T(dRaw1) + residual-branch gradient -> gB
      DRB_TRY(conv_backward(e, b.conv1, T1, xin, t1.d, t1.h, t1.w, gB, res_grad, nullptr, nullptr, nullptr, nullptr, s));
      std::swap(gA, gB);
    }
    if (li > 0) {
      // the stage input is the previous stage's output c_{li+1}, which also feeds lateral li
      e->launches += 1;
      DRB_TRY(drb_add_inplace(gA, e->dF[li], feats[li]->numel(), s));
    }
  }
  // ---- stem: x0 = maxpool(c1), c1 = relu(bn1(conv1(rgba))) ----
  e->launches += 2;
  DRB_TRY(drb_maxpool3d_backward(e->c1.f, gA, kG, e->c1.d, e->c1.h, e->c1.w, 64, gB, s));
  DRB_TRY(drb_add_inplace(gB, e->dF[0], e->c1.numel(), s));
  DRB_TRY(bn_backward(e, e->bn1, gB, e->c1.m(), e->c1.f, 1, gB, s));
  {
    const ConvW& w = e->conv1;
    GradPlanes gp;
    DRB_TRY(split_grad(e, gB, (long long)kG * e->c1.m(), 64, e->gp_hi, e->gp_lo, gp, s));
    DRB_TRY(engine_stem_im2col(e, io, s));
    DRB_TRY(wgrad(e, w, gp, e->col_hi, e->col_lo, kG, e->c1.d, e->c1.h, e->c1.w, 64, w.kpad, 1, 4, 125, GRAD(e, w.p_w),
                  nullptr, nullptr, 1.f, s));
  }
  return 0;
}

}  // namespace drb

extern "C" int drb_engine_backward(drb_engine* e, const drb_pair_io* io, const drb_pair_out* out,
                                   const drb_pair_grad* grad, cudaStream_t s) {
  DRB_REQUIRE(e && io && out && grad, "drb_engine_backward: null argument");
  DRB_REQUIRE(e->grad_mode && e->graph_valid,
              "drb_engine_backward: no training graph (switch grad mode on before encode / decode; one backward per forward)");
  DRB_REQUIRE(out->src_feats && out->tgt_feats && out->src_kp && out->tgt_kp && out->src_corr && out->tgt_corr &&
                  out->src_overlap && out->tgt_overlap,
              "drb_engine_backward: the forward outputs are needed");
  DRB_REQUIRE(io->src_grid && io->tgt_grid && io->src_mask && io->tgt_mask, "drb_engine_backward: null forward input");
  DRB_REQUIRE(io->n_src_mask == e->n_src_mask && io->n_tgt_mask == e->n_tgt_mask,
              "drb_engine_backward: io does not match the recorded forward");
  DRB_TRY(tokens_backward(e, out, grad, s));
  DRB_TRY(points_backward(e, io, s));
  DRB_TRY(fpn_backward(e, io, s));
  e->graph_valid = false;
  return 0;
}
