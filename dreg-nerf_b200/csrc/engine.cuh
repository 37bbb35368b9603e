// Types of the whole-path engine (engine.cu: forward, engine_bwd.cu: backward).
#pragma once
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

struct Param {
  std::string name;
  long long numel = 0;
  float* ptr = nullptr;
  float* grad = nullptr;     // bound gradient storage (training)
  bool trainable = false;
};

struct ConvW {               // packed weight planes
  int p_w = -1, p_b = -1;    // param indices (bias optional)
  int cout = 0, cin = 0, k = 1, stride = 1;
  bool im2col = false;       // lowered through an explicit im2col buffer
  int kpad = 0;              // im2col K (multiple of 64)
  plane_t* hi = nullptr;     // forward layout: [tap][cout][cin] (or [cout][kpad])
  plane_t* lo = nullptr;
  plane_t* thi = nullptr;    // data-gradient layout: [taps-1-tap][cin][cout] (or [kpad][cout]); training only
  plane_t* tlo = nullptr;
  bool need_dgrad = true;    // the stem convolution's input needs no gradient
  float* slot = nullptr;     // device {bits of max|w|, 1 / pre-scale}: acc_scale_dev = slot + 1
  int pack_index = -1;
};

struct BnP {
  int p_w = -1, p_b = -1, p_rm = -1, p_rv = -1;
  int c = 0;
  float* scale = nullptr;    // [g][c] (training graph only; the forward derives them on the fly)
  float* shift = nullptr;
  float* mean = nullptr;
  float* rstd = nullptr;
  float* raw_keep = nullptr; // the convolution output this BatchNorm normalised (training graph only)
  long long raw_elems = 0;
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

// What one transformer layer keeps for its backward pass (grad mode); in inference mode every layer
// points at the same scratch.
struct TSave {
  float *x1 = nullptr, *x2 = nullptr;                    // residual stream after self / cross attention
  plane_t *xn1_hi = nullptr, *xn1_lo = nullptr;          // LN1(x) + pos, LN2(x) + pos, LN3(x)
  plane_t *xn2_hi = nullptr, *xn2_lo = nullptr;
  plane_t *xn3_hi = nullptr, *xn3_lo = nullptr;
  float *qkv_s = nullptr, *qkv_c = nullptr;              // in_proj outputs [m][768]
  plane_t *att_s_hi = nullptr, *att_s_lo = nullptr;      // attention outputs (inputs of out_proj)
  plane_t *att_c_hi = nullptr, *att_c_lo = nullptr;
  plane_t *ffn_hi = nullptr, *ffn_lo = nullptr;          // relu(linear1)
};

// Device-side description of one weight for the multi-tensor pack kernel.
struct PackDesc {
  const float* w;
  int cout, cin, taps, kpad;     // kpad > 0: im2col layout
  plane_t *hi, *lo, *thi, *tlo;
  float* slot;
  long long fwd_elems, bwd_elems;
};

struct DsRound {             // one recorded round of the voxel-average down-sampling (grad mode)
  int n_in = 0, n_seg = 0;
  int* sorted_rows = nullptr;
  int* seg_start = nullptr;
};

}  // namespace drb

struct drb_engine {
  drb_engine_config cfg;
  std::vector<drb::Param> params;
  std::vector<void*> allocs;
  long long launches = 0;
  bool committed = false;
  bool profile = false;
  struct ProfRec { cudaEvent_t a, b; double flops; int list_id; double flops_per_tile; int m, cin, cout, k; };
  std::vector<ProfRec> prof;
  long long prof_forwards = 0;      // encode calls since the last profile read (tile-list averages)
  std::string fail;
  int max_tokens = 3000;

  // topology
  drb::ConvW conv1; drb::BnP bn1;
  std::vector<drb::Block> blocks[4];
  drb::ConvW pyr[5], ups[4];
  drb::TLayer tl[drb::kLayers];
  int fin_w, fin_b;                 // transformer_encoder.norm
  drb::ConvW q_proj, k_proj;
  int conf_w, conf_b;
  std::vector<drb::ConvW*> all_convs;       // every packed weight, in commit order
  drb::PackDesc* d_pack = nullptr;          // device table for the multi-tensor pack
  float* slots = nullptr;                   // device scale slots (2 floats per weight + gradient ring)
  int n_slots = 0, grad_slot_next = 0;

  // FPN buffers
  int D, H, W;                      // conv volume axes: D = Z, H = X, W = Y
  drb::plane_t *col_hi = nullptr, *col_lo = nullptr;   // shared im2col scratch
  long long col_elems = 0;
  float* raw = nullptr;             // shared raw conv output scratch (largest BN'd conv)
  long long raw_elems = 0;
  float* raw2 = nullptr;            // second scratch (downsample branch)
  double* bn_accum = nullptr;
  void* splitk_ws = nullptr; size_t splitk_ws_bytes = 0;   // split-K slices of the deep, few-tile GEMMs
  drb::Act c1, x0, c[4];            // c[0..3] = c2..c5
  std::vector<drb::Act> tmp;        // per-block temporaries
  drb::Act lat[5], sum[4], p[5];    // p[0] = p1 ... p[4] = p5
  // output-sparse evaluation of the two level-1 FPN convolutions
  bool sparse_fpn = true;
  bool fuse_topdown = true;     // FPN top-down merge in the lateral convolution's epilogue (DRB_FUSE_TOPDOWN=0: separate pass)
  bool bn_small = true;         // one-launch BatchNorm for the deep stages (DRB_BN_SMALL=0: the general path)
  bool update_running = true;   // training-mode BatchNorm writes running_mean / running_var (primary engine only)
  uint8_t* need = nullptr;
  int *tiles_out = nullptr, *tiles_in = nullptr, *tiles_in2 = nullptr, *tile_counts = nullptr;   // counts: int[3]
  unsigned long long* tile_totals = nullptr;   // running sums of the 3 list lengths (profiling)
  // point stage
  float* rows = nullptr;            // [2*max_mask][260]
  float* rows_ds = nullptr;
  void* ds_ws = nullptr; size_t ds_ws_bytes = 0;
  int n_src = 0, n_tgt = 0;         // tokens after down-sampling
  int n_src_mask = 0, n_tgt_mask = 0;
  std::vector<drb::DsRound> ds_tape;
  int* ds_tape_buf = nullptr; long long ds_tape_cap = 0;
  // transformer buffers (capacity tok_cap rows)
  int tok_cap = 0;
  bool tok_grad = false;            // token buffers were allocated with the training extras
  float *x = nullptr, *pos = nullptr, *qkv = nullptr, *kp_xyz = nullptr, *sbuf = nullptr;
  drb::plane_t *xn_hi = nullptr, *xn_lo = nullptr, *att_hi = nullptr, *att_lo = nullptr;
  drb::plane_t *ffn_hi = nullptr, *ffn_lo = nullptr, *dec_hi = nullptr, *dec_lo = nullptr;
  drb::plane_t *qp_hi = nullptr, *qp_lo = nullptr, *kp_hi = nullptr, *kp_lo = nullptr;
  std::vector<void*> tok_allocs;
  bool tc_attention = false;        // attention through the tcgen05 kernel (attention.cu)
  void* att_ws = nullptr; size_t att_ws_bytes = 0;

  // ---- training graph ----
  bool grad_mode = false;
  bool graph_valid = false;         // a grad-mode forward has run and its buffers are intact
  bool grad_allocated = false;
  float* xs[drb::kLayers + 1] = {}; // residual stream at the start of every layer (+ final)
  drb::TSave ts[drb::kLayers];
  float *qf = nullptr, *kf = nullptr;               // decoder projections in fp32 [6 m][256]
  // backward scratch (FPN part, sized by the largest activation)
  long long g_elems = 0;
  float *gA = nullptr, *gB = nullptr, *gT1 = nullptr, *gT2 = nullptr;
  drb::plane_t *gp_hi = nullptr, *gp_lo = nullptr;
  float* dF[5] = {};                // gradients of the backbone features c1..c5 from the lateral convolutions
  float* dcol = nullptr; long long dcol_elems = 0;
  float* wg_stage = nullptr; long long wg_stage_elems = 0;   // weight-gradient tile staging ([Cout][taps * Cin])
  double* bn_sums = nullptr;
  // backward scratch (token part)
  float *t_dx = nullptr, *t_dy = nullptr, *t_dh = nullptr, *t_dqkv = nullptr, *t_datt = nullptr, *t_dxn = nullptr;
  float *t_ddec = nullptr, *t_dqf = nullptr, *t_dkf = nullptr;
  float *t_dcorr = nullptr, *t_dov = nullptr;       // [6][m][3], [6][m]
  float* t_drows = nullptr;                         // [2*max_mask][256] (two ping-pong halves live in gT1/gT2)
  drb::plane_t *t_ph = nullptr, *t_pl = nullptr;    // plane scratch [6 m][1024]
  void* mha_ws = nullptr; size_t mha_ws_bytes = 0;

  template <typename T> T* alloc(long long n) {
    void* q = nullptr;
    if (n <= 0) n = 1;
    if (cudaMalloc(&q, (size_t)n * sizeof(T)) != cudaSuccess) {
      fail = "cudaMalloc failed";
      return nullptr;
    }
    allocs.push_back(q);
    return (T*)q;
  }
  int add_param(const std::string& name, long long numel, bool trainable) {
    drb::Param q; q.name = name; q.numel = numel; q.trainable = trainable;
    params.push_back(q);
    return (int)params.size() - 1;
  }
};

namespace drb {

#define DRB_TRY(expr)            \
  do {                           \
    int _rc = (expr);            \
    if (_rc != 0) return _rc;    \
  } while (0)

static inline float* P(drb_engine* e, int idx) { return idx >= 0 ? e->params[idx].ptr : nullptr; }
static inline float* GRAD(drb_engine* e, int idx) { return idx >= 0 ? e->params[idx].grad : nullptr; }
static inline plane_t* off(plane_t* p, long long n) { return p ? p + n : nullptr; }

// engine.cu
int engine_list_id(const drb_engine* e, const int* tile_list);
int engine_run_igemm(drb_engine* e, const ConvW& w, const plane_t* x_hi, const plane_t* x_lo, int g, int d, int h,
                     int wd, int cin, int k, const float* bias, const float* residual, int relu, float scale,
                     float* out, plane_t* out_hi, plane_t* out_lo, long long ld, cudaStream_t s,
                     const plane_t* w_hi = nullptr, const plane_t* w_lo = nullptr, int cout_override = 0,
                     const int* tile_list = nullptr, const int* tile_count = nullptr,
                     const float* scale_dev0 = nullptr, const float* scale_dev1 = nullptr,
                     const int* res_dims = nullptr);      // {d, h, w} of a coarser residual added with nearest x2 up-sampling
int engine_ensure_tokens(drb_engine* e, int m);
int engine_stem_im2col(drb_engine* e, const drb_pair_io* io, cudaStream_t s);
int engine_im2col(drb_engine* e, const ConvW& w, const Act& in, cudaStream_t s);
// engine_bwd.cu
int engine_ensure_grad_buffers(drb_engine* e);

}  // namespace drb
