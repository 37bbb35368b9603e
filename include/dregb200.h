/*
 * libdregb200 - C ABI of the B200-native DReg-NeRF registration hot path.
 *
 * Plain pointers and sizes only: every pointer is a CUDA device pointer unless the name says
 * "host"; every call is asynchronous on `stream` unless documented otherwise; every function
 * returns DRB_OK (0) or a negative DRB_E* code and records a message readable through
 * drb_last_error().  No exceptions cross the boundary, no torch types, no hidden global state
 * other than a lazily created device error flag and the cached driver entry point.
 *
 * 16-bit "planes": an fp32 tensor x is carried either as ONE bf16 plane (hi only; the bf16
 * configurations) or as a PAIR of fp16 planes hi = fp16(x), lo = fp16(x - hi) (22 significand bits;
 * fp32-grade GEMM results at three tensor-core products).  Everywhere below a NULL lo pointer selects
 * the single bf16 plane and a non-NULL lo pointer selects the fp16 pair.  |x| > 65504 saturates in
 * pair mode; weights are pre-scaled by a power of two (drb_weight_scale) to sit in fp16's normal range.
 * Activations are channels-last: [g][d][h][w][c] with d = Z, h = X, w = Y of the reference's
 * [1, C, Z, X, Y] tensors (conerf/register/nerf_regtr.py:112-134).
 *
 * Each entry point cites the reference interface it replaces (paths relative to the DReg-NeRF
 * repository root).
 */
#ifndef DREGB200_H_
#define DREGB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DRB_OK 0
#define DRB_EINVAL (-1)
#define DRB_ECUDA (-2)
#define DRB_ENOMEM (-3)
#define DRB_ESTATE (-4)

#define DRB_ABI_VERSION 4

typedef struct CUstream_st* drb_stream_t; /* == cudaStream_t */

int drb_abi_version(void);
const char* drb_last_error(void);
/* Value of the device-side pipeline watchdog flag (0 = healthy).  Synchronous. */
int drb_igemm_error_flag(int* host_value);
/* The flag of the current device is sticky; codes: 1-4 / 11-14 tcgen05 pipeline watchdogs (igemm / wgrad), 21 mask
 * index out of range, 31 surface-field marcher watchdog (result truncated).  Clears it. */
int drb_error_flag_clear(void);
/* The flag as the host sees it NOW, without waiting for the device (it lives in mapped pinned host memory): valid
 * for work the caller has already synchronised with - e.g. after cudaStreamSynchronize of its own stream, while
 * other streams keep running (drb_igemm_error_flag waits for the whole device). */
int drb_error_flag_peek(int* host_value);
/* Debug: the flag and 15 per-site detail words (which watchdog sites fired).  Synchronises. */
int drb_error_flag_detail(int* host16);

/* ------------------------------------------------------------------------------------------
 * R2 / R6 / R7: nn.Conv3d (conerf/model/resnet3d.py:81-86,120, feature_pyramid_net.py:24,33)
 * and nn.Linear / MultiheadAttention in/out projections (conerf/register/transformer.py:128-138,
 * nerf_regtr.py:268-270) as one tcgen05 implicit GEMM.  Stride 1, padding k/2.
 *   out = relu?((conv(x, w) * acc_scale + bias) * out_scale + residual)
 * A linear layer over n tokens is the 1x1x1 case with g = d = h = 1, w = n.
 * ---------------------------------------------------------------------------------------- */
typedef struct drb_conv3d_desc {
  int g, d, h, w;          /* activation volume (output == input extent)                    */
  int cin, cout;           /* cin multiple of 64                                            */
  int kd, kh, kw;          /* odd kernel extents                                            */
  int planes;              /* 1: one bf16 plane, 2: fp16 hi/lo pair (fp32-grade)            */
  int relu;
  float acc_scale;         /* 1 / (weight pre-scale); 0 is read as 1                        */
  float out_scale;         /* 0 is read as 1                                                */
  const void* x_hi;        /* 16-bit [g][d][h][w][cin]                                      */
  const void* x_lo;        /* same shape, planes == 2 only                                  */
  const void* w_hi;        /* 16-bit [kd*kh*kw][cout][cin]  (see drb_pack_conv_weight)      */
  const void* w_lo;
  const float* bias;       /* [cout] or NULL                                                */
  const float* residual;   /* fp32 [m][ld_out] or NULL                                      */
  float* out;              /* fp32 [m][ld_out] or NULL,  m = g*d*h*w                        */
  void* out_hi;            /* 16-bit [m][ld_out] or NULL                                    */
  void* out_lo;            /* fp16 [m][ld_out] or NULL (non-NULL => hi is fp16 too)         */
  long long ld_out;        /* row pitch in elements, 0 -> cout; multiple of 8               */
  /* Optional fused BatchNorm statistics: double [g][cout][2] receiving per-channel sum and sum of
   * squares of the fp32 output (zeroed by the call; plain epilogue only: no bias/residual/relu).      */
  double* bn_accum;
  /* Output-sparse mode (optional, both or neither): only the 128-row output tiles listed in
   * tile_list[0 .. *tile_count) are computed; rows of other tiles are left untouched.  Tile
   * numbering: see drb_conv3d_tile_shape.  Device pointers.                                         */
  const int* tile_list;
  const int* tile_count;
  /* Optional device scalars multiplied into acc_scale inside the kernel (pre-scales that only exist on the
   * device: packed weights, gradient planes of drb_grad_split); NULL = 1.                               */
  const float* acc_scale_dev[2];
  /* Optional split-K workspace (device).  Few output tiles but a long reduction (deep backbone layers) are split
   * over K: every slice stores its partial tile here and a second kernel adds the slices in index order, so the
   * result is run-to-run bit-stable (no atomics).  NULL / too small: the reduction is not split.            */
  void* splitk_ws;
  size_t splitk_ws_bytes;
  /* Optional: `residual` is a COARSER volume [g][res_d][res_h][res_w][ld_out] added with nearest x2 up-sampling
   * (row (g, d, h, w) takes residual row (g, d / 2, h / 2, w / 2)): the FPN's top-down merge
   * F.interpolate(top, scale_factor=2)[..., :d, :h, :w] + lateral (feature_pyramid_net.py:58-61) fused into the
   * lateral convolution's epilogue.  All zero: residual has the output's own shape.                       */
  int res_d, res_h, res_w;
} drb_conv3d_desc;
/* Geometry of the 128-row output tiles drb_conv3d_igemm uses for a [g][d][h][w] volume: box extents
 * (bg, bd, bh, bw) and tile counts (tg, td, th, tw); tile index = ((ig*td + id)*th + ih)*tw + iw. */
int drb_conv3d_tile_shape(int g, int d, int h, int w, int box[4], int tiles[4]);
int drb_conv3d_igemm(const drb_conv3d_desc* desc, drb_stream_t stream);

/* fp32 -> planes (lo == NULL: one bf16 plane; else fp16 hi/lo pair). */
int drb_split_planes(const float* x, void* hi, void* lo, long long n, drb_stream_t stream);

/* Power-of-two pre-scale that maps max|w| into [256, 512) (fp16 normal range for hi AND lo).
 * Synchronous: reads 4 bytes back.  Pass 1 / scale as drb_conv3d_desc.acc_scale. */
int drb_weight_scale(const float* w, long long n, float* host_scale, drb_stream_t stream);
/* torch Conv3d weight [cout][cin][taps] fp32, times `scale` -> planes [taps][cout][cin_pad], zero
 * padded. */
int drb_pack_conv_weight(const float* w, int cout, int cin, int taps, int cin_pad, float scale,
                         void* hi, void* lo, drb_stream_t stream);
/* same weight -> planes [cout][kpad] with k = tap*cin + c (pairs with drb_im2col). */
int drb_pack_conv_weight_im2col(const float* w, int cout, int cin, int taps, int kpad, float scale,
                                void* hi, void* lo, drb_stream_t stream);

/* Generic strided / large-kernel convolutions (conv1 5^3 s2, resnet3d.py:120; 3^3 s2 and 1^3 s2
 * in the first Bottleneck of layer2-4, resnet3d.py:83,140-146) are lowered to a 1x1x1 igemm over
 * an explicit im2col buffer: planes [g][do][ho][wo][kpad], k = ((kz*K + ky)*K + kx)*c + ch. */
typedef struct drb_im2col_desc {
  const float* x;                 /* fp32, arbitrary element strides                        */
  long long sg, sc, sd, sh, sw;
  int g, c, d, h, w;              /* input extent                                           */
  int k, stride, pad;             /* cubic kernel                                           */
  int kpad;                       /* >= k^3*c, multiple of 64                               */
} drb_im2col_desc;
int drb_im2col(const drb_im2col_desc* desc, void* hi, void* lo, drb_stream_t stream);
/* Stem fast path (conv1: 5^3, stride 2, pad 2 over 4 channels, resnet3d.py:120): x points at the first
 * of 4 channels of ONE grid with element strides (sc, sd, sh, sw); scratch holds d*h*w*4 floats;
 * planes [do][ho][wo][512] with the k order of drb_im2col.  Same result as drb_im2col. */
int drb_im2col_stem(const float* x, long long sc, long long sd, long long sh, long long sw, int d, int h,
                    int w, void* scratch, void* hi, void* lo, drb_stream_t stream);

/* nn.BatchNorm3d (resnet3d.py:82-87,121).  x is fp32 [g][m][c]; statistics are per (g, c):
 * the reference runs one grid per call with batch size 1 (nerf_regtr.py:135).  accum is
 * [g][c][2] doubles of scratch. */
int drb_bn_stats(const float* x, int g, long long m, int c, double* accum, drb_stream_t stream);
/* training != 0: batch statistics, running buffers updated sequentially over g with momentum
 * (unbiased variance), as g successive forward calls would.  training == 0: running statistics.
 * Writes scale/shift [g][c] such that y = x*scale + shift. */
int drb_bn_finalize(const double* accum, int g, long long m, int c, const float* gamma,
                    const float* beta, float* running_mean, float* running_var, int training,
                    float momentum, float eps, float* scale, float* shift, drb_stream_t stream);
/* y = relu?(x*scale[g][c] + shift[g][c] + residual); scale/shift may be NULL (identity). */
int drb_scale_shift_act(const float* x, const float* scale, const float* shift,
                        const float* residual, int relu, int g, long long m, int c, float* out,
                        void* out_hi, void* out_lo, drb_stream_t stream);

/* drb_bn_finalize + drb_scale_shift_act in one launch (what the engine uses): statistics from `accum`
 * (training) or the running buffers (eval), running-statistics update, normalise, residual, ReLU. */
int drb_bn_apply(const float* x, const double* accum, int g, long long m, int c, const float* gamma,
                 const float* beta, float* running_mean, float* running_var, int training, float momentum,
                 float eps, const float* residual, int relu, float* out, void* out_hi, void* out_lo,
                 drb_stream_t stream);

/* The same BatchNorm (nn.BatchNorm3d of resnet3d.py:88-100 in both modes) for SMALL tensors - the deep ResNet
 * stages, m <= 1024 rows per grid - in one launch: fp64 statistics, running-buffer update, normalise (+ residual,
 * ReLU, plane split) and, when the save buffers are given ([g][c] each; all four or none), the mean / rstd /
 * scale / shift the backward pass needs (drb_bn_save_stats).  drb_bn_small_supported tells whether (m, c) qualifies. */
int drb_bn_small_supported(long long m, int c);
int drb_bn_small(const float* x, int g, long long m, int c, const float* gamma, const float* beta,
                 float* running_mean, float* running_var, int training, float momentum, float eps,
                 const float* residual, int relu, float* out, void* out_hi, void* out_lo, float* save_mean,
                 float* save_rstd, float* save_scale, float* save_shift, drb_stream_t stream);

/* nn.MaxPool3d(3, 2, 1) (resnet3d.py:123) on fp32 [g][d][h][w][c]. */
int drb_maxpool3d(const float* x, int g, int d, int h, int w, int c, float* out, void* out_hi,
                  void* out_lo, drb_stream_t stream);
/* FeaturePyramid_v1._upsample + torch.add (feature_pyramid_net.py:58-61,73-103):
 * out[d][h][w] = coarse[d/2][h/2][w/2] + lateral[d][h][w]. */
int drb_upsample2_add(const float* coarse, int dc, int hc, int wc, const float* lateral, int g,
                      int d, int h, int w, int c, float* out, void* out_hi, void* out_lo,
                      drb_stream_t stream);

/* R3: F.interpolate(trilinear, align_corners=True) evaluated only at the masked voxels, plus the
 * xyz gather (nerf_regtr.py:138-147).  p1 is one grid [dc][hc][wc][c]; mask holds flat indices
 * x*(Y*Z) + y*Z + z; rows_out[i] = [x y z 0 | c features], pitch ld_rows (>= 4 + c). */
int drb_trilinear_gather(const float* p1, int dc, int hc, int wc, int c, const float* grid,
                         long long s_ch, long long s_z, long long s_x, long long s_y, int X, int Y,
                         int Z, const long long* mask, int k, float* rows_out, int ld_rows,
                         drb_stream_t stream);

/* Output-sparse FPN: p1 is only read by drb_trilinear_gather, so pyramid_transformation_1 and
 * upsample_transform_1 (84 % of the FLOPs) are evaluated on the tiles the gather needs.  need: uint8
 * [g][dc][hc][wc] scratch; masks_host / ks_host: host arrays of g device mask pointers and lengths;
 * list_out: tiles holding a needed p1 voxel; list_in: tiles holding a voxel within one voxel of one
 * (inputs of the 3^3 convolution); counts: device int[2].  Tile numbering: drb_conv3d_tile_shape. */
int drb_fpn_need_tiles(const long long* const* masks_host, const int* ks_host, int g, int X, int Y, int Z,
                       int dc, int hc, int wc, uint8_t* need, int* list_out, int* list_in, int* counts,
                       unsigned long long* totals /* optional running totals [2] */, drb_stream_t stream);

/* Tiles within `dilate` voxels of a voxel marked by drb_fpn_need_tiles (backward pass: the data gradient of
 * a 3^3 convolution spreads one voxel per layer).  count: device int, zeroed by the call. */
int drb_fpn_dilated_tiles(const uint8_t* need, int g, int dc, int hc, int wc, int dilate, int* list, int* count,
                          unsigned long long* total /* optional running total */, drb_stream_t stream);

/* R4: hierarchical voxel-average down-sampling (conerf/register/grid_downsample.py:6-94).
 * rows [n_src + n_tgt][ld] = [x y z 0 | 256 features]; cells of size dl0 * 2^round; rows of one
 * cloud are emitted in ascending (cx, cy, cz) order; stops after the first round that leaves
 * <= max_total rows.  Synchronous (reads the counts back).  workspace from
 * drb_downsample_workspace_bytes(n). rows_out may alias neither input nor workspace. */
size_t drb_downsample_workspace_bytes(int n_rows, int ld);
int drb_hierarchical_downsample(const float* rows, int n_src, int n_tgt, int ld, int num_rounds,
                                double dl0, int max_total, void* workspace, size_t workspace_bytes,
                                float* rows_out, int* host_n_src_out, int* host_n_tgt_out,
                                drb_stream_t stream);

/* Same, recording every round for the backward pass: tape receives, per round, the sort permutation
 * (n_in ints) followed by the segment starts (n_seg + 1 ints); host_round_info[2 r] = n_in,
 * [2 r + 1] = n_seg; *host_rounds = rounds run.  Replay in reverse with drb_segment_mean_backward. */
int drb_hierarchical_downsample_tape(const float* rows, int n_src, int n_tgt, int ld, int num_rounds,
                                     double dl0, int max_total, void* workspace, size_t workspace_bytes,
                                     float* rows_out, int* host_n_src_out, int* host_n_tgt_out, int* tape,
                                     long long tape_capacity, int* host_round_info, int* host_rounds,
                                     drb_stream_t stream);

/* R5: PositionEmbeddingCoordsSine (conerf/register/position_embedding.py:30-53), d_model 256. */
int drb_pos_embed_sine(const float* xyz, int ld_xyz, int n, float scale, float* out,
                       drb_stream_t stream);
/* nn.LayerNorm(256) (+ optional positional add) (transformer.py:237-238,262-264,285,291).
 * out / out_hi / out_lo receive LN(x) + add; any of them may be NULL. */
int drb_layernorm256(const float* x, int n, const float* gamma, const float* beta,
                     const float* add, float* out, void* out_hi, void* out_lo,
                     drb_stream_t stream);
/* softmax(q k^T * scale) v per head (torch nn.MultiheadAttention core, transformer.py:240-281). */
int drb_mha_core(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv, int nq,
                 int nk, int heads, float scale, float* out, void* out_hi, void* out_lo, int ld_out,
                 drb_stream_t stream);
/* The same attention core on the 5th-gen tensor cores (tcgen05 + TMEM + TMA, FlashAttention style: logits and
 * probabilities never leave the SM).  drb_mha_tc_pack converts the fp32 q / k / v rows of ONE in_proj output
 * ([n][ld], head h at columns 32 h ..) into per-head 16-bit planes inside `workspace` (1024-byte aligned,
 * drb_mha_tc_workspace_bytes); the rows form two segments - the source cloud [0, split) and the target cloud
 * [split, n) (split = n: one segment).  drb_mha_tc_forward attends the queries of segment q_seg (-1: both
 * segments in one launch) to the keys / values of their own segment (cross == 0, self-attention) or of the other
 * one (cross != 0): a transformer layer is two packs and two launches.  Output rows are the queries' input row
 * indices, columns 32 h .. of out / out_hi / out_lo (pitch ld_out). */
size_t drb_mha_tc_workspace_bytes(int n, int heads, int planes);
int drb_mha_tc_pack(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv, int n, int split,
                    int heads, int planes, float scale, void* workspace, size_t workspace_bytes, drb_stream_t stream);
int drb_mha_tc_forward(const void* workspace, int n, int split, int heads, int planes, int q_seg, int cross,
                       float* out, void* out_hi, void* out_lo, int ld_out, drb_stream_t stream);
/* CorrespondenceDecoder.simple_attention tail (nerf_regtr.py:292-306): row softmax of s [nq][ld]
 * over nk keys, weighted sum of xyz [nk][ld_xyz] -> out [nq][3]. */
int drb_softmax_weighted_xyz(const float* s, int ld, int nq, int nk, const float* xyz, int ld_xyz,
                             float* out, drb_stream_t stream);
/* sigmoid(w . feat + b) (nerf_regtr.py:384-387); feat [n][256] -> out [n]. */
int drb_overlap_sigmoid(const float* feat, int n, const float* w, const float* b, float* out,
                        drb_stream_t stream);

/* R8 + R9: weighted Procrustes (conerf/register/se3.py:89-140) over the stacked correspondences
 * of nerf_regtr.py:208-230, for `layers` problems at once.  Segment 1 has n1 rows, segment 2 has
 * n2 rows; *_ls are the per-layer strides in floats (0 broadcasts).  out [layers][3][4]. */
int drb_procrustes(const float* a1, long long a1_ls, const float* b1, long long b1_ls,
                   const float* w1, long long w1_ls, int n1, const float* a2, long long a2_ls,
                   const float* b2, long long b2_ls, const float* w2, long long w2_ls, int n2,
                   int ld_pts, int layers, float* out, drb_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * A1-A6: extract (conerf/radiance_fields/ngp.py:148-193, conerf/register/sample_grid.py:208-343,
 * conerf/utils/nerfacc_utils.py:84-222, eval_ngp_nerf.py:337-412)
 * ---------------------------------------------------------------------------------------- */
typedef struct drb_ngp_params {
  const float* hash_table;   /* fp32 [total_entries][2], levels concatenated                 */
  const float* w1;           /* density MLP 32 -> 64, row major [64][32]                      */
  const float* w2;           /* 64 -> 16, [16][64]                                            */
  const float* c1;           /* colour MLP [64][32] (input = 16 SH | 15 feat | 0)             */
  const float* c2;           /* [64][64]                                                      */
  const float* c3;           /* [3][64] (tcnn pads to 16 rows; only 3 are used)               */
  float aabb[6];
} drb_ngp_params;
/* Number of [2]-float entries of the 16-level hash table (levels concatenated). Host only. */
long long drb_ngp_table_entries(void);
/* NGPradianceField.query_density (ngp.py:148-176): x world [n][3] -> density [n], feat [n][15]. */
int drb_ngp_density(const drb_ngp_params* p, const float* x, int n, float* density, float* feat,
                    drb_stream_t stream);
/* mean over ndirs fixed view directions of NGPradianceField.query_rgb (ngp.py:178-193,
 * sample_grid.py:332-340): feat [n][15], dirs host [ndirs][3] -> rgb [n][3]. */
int drb_ngp_rgb_mean(const drb_ngp_params* p, const float* feat, int n, const float* host_dirs,
                     int ndirs, float* rgb, drb_stream_t stream);
/* Surface-field mask (sample_grid.py:245-318 + nerfacc 0.3.5 ray marching): for every point p and
 * camera origin o march o->p through the binary occupancy grid with `step`; surface[p] = 1 when
 * max over cameras/samples of alpha*T >= cut_off. */
int drb_surface_mask(const drb_ngp_params* p, const uint8_t* occ_binary, int res,
                     const float* roi_aabb_host, const float* scene_aabb_host, const float* points,
                     int n, const float* cam_origins, int ncams, float step, float cut_off,
                     uint8_t* surface, drb_stream_t stream);
/* The same call with its scratch (Morton sort buffers, coarse bitmap, cell-major records of the dense hash levels)
 * in a caller workspace of at least drb_surface_mask_workspace_bytes(n) bytes: no allocation inside the call. */
size_t drb_surface_mask_workspace_bytes(int n);
int drb_surface_mask_ws(const drb_ngp_params* p, const uint8_t* occ_binary, int res,
                        const float* roi_aabb_host, const float* scene_aabb_host, const float* points,
                        int n, const float* cam_origins, int ncams, float step, float cut_off,
                        uint8_t* surface, void* workspace, size_t workspace_bytes, drb_stream_t stream);
/* SampleGrid sampling + Evaluator.sample_points scatter (sample_grid.py:223-243,
 * eval_ngp_nerf.py:383-412): fused extract of one NeRF block; see INTEGRATION.md. */
typedef struct drb_extract_desc {
  int res;                       /* grid resolution R                                          */
  float roi_aabb[6];
  float scene_aabb[6];
  const long long* occupied;     /* flat indices of occupied cells (nonzero(binary))           */
  int n_occupied;
  const float* jitter;           /* U[0,1) [n_occupied][3] (torch.rand_like in the reference)   */
  const uint8_t* occ_binary;     /* [R][R][R]                                                  */
  const float* cam_origins;      /* [ncams][3]                                                 */
  int ncams;
  float render_step_size;
  float density_thre;            /* 0.7 */
  float cut_off;                 /* 0.5 */
  const float* host_dirs;        /* [ndirs][3]                                                 */
  int ndirs;
  int surface_only_where_dense;  /* 1: march rays only for cells with density > density_thre; their
                                  * surface_mask entries stay 0.  Exact for voxel_grid / voxel_mask
                                  * (eval_ngp_nerf.py:383 keeps surface & density only).             */
  int rgb_only_where_masked;     /* 1: evaluate the colour head only for cells that pass both masks (after the
                                  * march); rgb rows of the other cells are 0.  Exact for voxel_grid.       */
} drb_extract_desc;
/* Writes points [n][3], rgb [n][3], alpha [n], density_mask [n], surface_mask [n] and scatters
 * rows of cells with both masks set into voxel_grid [R^3][7] (pre-zeroed by the call). */
int drb_extract_block(const drb_ngp_params* p, const drb_extract_desc* e, float* points, float* rgb,
                      float* alpha, uint8_t* density_mask, uint8_t* surface_mask, float* voxel_grid,
                      drb_stream_t stream);

/* drb_extract_block with all scratch (features, densities, index list, the marcher's buffers) in a caller
 * workspace of at least drb_extract_workspace_bytes(n_occupied) bytes: no allocation inside the call
 * (drb_extract_block itself uses the stream-ordered allocator). */
size_t drb_extract_workspace_bytes(int n_occupied);
int drb_extract_block_ws(const drb_ngp_params* p, const drb_extract_desc* e, float* points, float* rgb,
                         float* alpha, uint8_t* density_mask, uint8_t* surface_mask, float* voxel_grid,
                         void* workspace, size_t workspace_bytes, drb_stream_t stream);

/* Device time (ms) of the surface-field kernel inside the calling thread's most recent
 * drb_extract_block (roofline instrumentation; waits for that kernel). */
int drb_extract_last_surface_ms(float* host_ms);

/* Profile mode (per calling thread): while on, every drb_extract_block records the surface-field
 * kernel's start/stop events; drb_extract_read_profile waits for them, returns the summed kernel time
 * and the number of launches, and clears the record. */
int drb_extract_set_profile(int on);
int drb_extract_read_profile(float* total_ms, int* launches);

/* Work counters of the surface-field kernel accumulated on the current device since the last reset:
 * host4 = {rays marched, empty-space skip events, density samples, warp scheduling rounds}.
 * Synchronises the device (roofline instrumentation: issued gather bytes = samples * 1024). */
int drb_march_stats(unsigned long long* host4, int reset);

/* ------------------------------------------------------------------------------------------
 * Backward pass (train_nerf_regtr.py:229 `loss.backward()`; the reference relies on torch autograd
 * through the modules cited next to each forward entry point above).
 * ---------------------------------------------------------------------------------------- */
/* Weight gradient of a stride-1 "same" Conv3d / Linear: dw[co][c][tap] += scale * sum_m dy[m][co] *
 * x[m + off(tap)][ci] with k = tap * cin + ci re-read as (tap', c) = (k / c_real, k % c_real) - identity
 * for a direct convolution, the im2col unpacking when x is an im2col buffer (kd = kh = kw = 1,
 * cin = kpad, c_real = channels, taps_real = K^3).  Both operands are 16-bit planes [g][d][h][w][c]
 * (tcgen05, MN-major operands).  tile_list: reduce only over the listed 128-voxel tiles
 * (numbering of drb_conv3d_tile_shape). */
typedef struct drb_wgrad_desc {
  int g, d, h, w;
  int cout, cin;             /* cin multiple of 64, cout multiple of 8                          */
  int kd, kh, kw;
  int planes;
  const void* dy_hi; const void* dy_lo;    /* [g][d][h][w][cout]                               */
  const void* x_hi; const void* x_lo;      /* [g][d][h][w][cin]                                */
  float scale;                             /* 0 is read as 1                                   */
  const float* scale_dev;                  /* optional device scalar multiplied into scale     */
  float* dw;                               /* fp32 [cout][c_real][taps_real], accumulated      */
  int c_real, taps_real;                   /* 0 -> cin, kd*kh*kw                               */
  const int* tile_list;
  const int* tile_count;
  float* stage;                            /* optional workspace, >= cout * kd*kh*kw * cin floats: the    */
  long long stage_elems;                   /* tile is reduced in GEMM layout and transposed into dw;      */
                                           /* NULL: scattered atomics straight into dw (slow for k > 1)   */
} drb_wgrad_desc;
int drb_conv3d_wgrad(const drb_wgrad_desc* desc, drb_stream_t stream);

/* fp32 [rows][cols] (pitch ld_in) -> planes (pitch ld_out, padding columns zeroed).  Pair mode (lo != NULL):
 * the tensor is multiplied by the power of two that maps max|x| into [256, 512) first; slot (2 device
 * floats of scratch) receives {bits of max|x|, 1 / scale} - pass slot + 1 as acc_scale_dev / scale_dev. */
int drb_grad_split(const float* x, long long rows, int cols, long long ld_in, long long ld_out, void* hi,
                   void* lo, float* slot, drb_stream_t stream);
int drb_add_inplace(float* dst, const float* src, long long n, drb_stream_t stream);
/* dy[i] = 0 where the saved ReLU output plane is zero. */
int drb_relu_mask_plane(float* dy, const void* hi, long long n, drb_stream_t stream);
/* out[c] += sum_r x[r][c]  (bias gradients). */
int drb_colsum_add(const float* x, long long rows, int c, long long ld, float* out, drb_stream_t stream);
/* BatchNorm3d: per-(g, c) mean / rstd / scale / shift exactly as drb_bn_apply derives them. */
int drb_bn_save_stats(const double* accum, int g, long long m, int c, const float* gamma, const float* beta,
                      const float* running_mean, const float* running_var, int training, float eps,
                      float* mean, float* rstd, float* scale, float* shift, drb_stream_t stream);
/* BatchNorm3d (+ fused ReLU) backward.  dy [g][m][c] is masked IN PLACE when relu != 0 (post > 0 if post is
 * given, else fma(raw, scale, shift) > 0); dx may alias dy; sums: [g][c][2] doubles of scratch;
 * dgamma / dbeta (optional) are accumulated. */
int drb_bn_backward(float* dy, const float* raw, const float* post, const float* scale, const float* shift,
                    const float* mean, const float* rstd, const float* gamma, int relu, int training, int g,
                    long long m, int c, double* sums, float* dx, float* dgamma, float* dbeta,
                    drb_stream_t stream);
int drb_maxpool3d_backward(const float* x, const float* dout, int g, int d, int h, int w, int c, float* dx,
                           drb_stream_t stream);
int drb_upsample2_add_backward(const float* dsum, int g, int d, int h, int w, int c, int dc, int hc, int wc,
                               float* dcoarse, drb_stream_t stream);
/* drows [k][ld] with the feature gradient at columns [col0, col0 + c); dp1 must be zeroed by the caller. */
int drb_trilinear_gather_backward(const float* drows, int ld, int col0, int dc, int hc, int wc, int c, int X,
                                  int Y, int Z, const long long* mask, int k, float* dp1, drb_stream_t stream);
/* Adjoint of drb_im2col: dx[g][d][h][w][c] = residual + sum over taps of dcol. */
int drb_col2im(const float* dcol, int g, int c, int d, int h, int w, int k, int stride, int pad, int kpad,
               const float* residual, float* dx, drb_stream_t stream);
/* dx (+)= dLayerNorm/dx; dgamma / dbeta (optional) accumulated. */
int drb_layernorm256_backward(const float* x, const float* dy, int n, const float* gamma, float* dx,
                              int overwrite, float* dgamma, float* dbeta, drb_stream_t stream);
int drb_overlap_sigmoid_backward(const float* feat, const float* ov, const float* dov, int n, const float* w,
                                 float* dfeat, float* dw, float* db, drb_stream_t stream);
int drb_softmax_rows(float* s, long long rows, int nk, int ld, float scale, drb_stream_t stream);
int drb_softmax_backward_rows(const float* p, float* dp, long long rows, int nk, int ld, float scale,
                              drb_stream_t stream);
/* s [nq][ld] (logits) -> dS of corr = softmax(s) xyz given dcorr [nq][3], in place. */
int drb_softmax_weighted_xyz_backward(float* s, int ld, int nq, int nk, const float* xyz, int ld_xyz,
                                      const float* dcorr, drb_stream_t stream);
/* C[b](m, n) (+)= alpha sum_k A[b](m, k) B[b](k, n), arbitrary element strides, fp32 CUDA cores. */
int drb_sgemm_strided(const float* A, long long a_b, long long a_m, long long a_k, const float* B,
                      long long b_b, long long b_k, long long b_n, float* C, long long c_b, long long c_m, int M,
                      int N, int K, int batch, float alpha, int accumulate, drb_stream_t stream);
size_t drb_mha_backward_workspace_bytes(int nq, int nk, int heads);
int drb_mha_core_backward(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv,
                          const float* dout, int ld_do, int nq, int nk, int heads, float scale, float* dq,
                          int ld_dq, float* dk, int ld_dk, float* dv, int ld_dv, void* workspace,
                          size_t workspace_bytes, drb_stream_t stream);
/* Gradients of drb_procrustes w.r.t. its point / weight operands (accumulated; NULL = not needed). */
int drb_procrustes_backward(const float* a1, long long a1_ls, const float* b1, long long b1_ls,
                            const float* w1, long long w1_ls, int n1, const float* a2, long long a2_ls,
                            const float* b2, long long b2_ls, const float* w2, long long w2_ls, int n2,
                            int ld_pts, int layers, const float* dpose, float* da1, float* db1, float* dw1,
                            float* da2, float* db2, float* dw2, drb_stream_t stream);
int drb_segment_mean_backward(const float* dout, int ld_out, const int* sorted_rows, const int* seg_start,
                              int n_seg, int c, float* din, int ld_in, drb_stream_t stream);

/* Fused clip_grad_norm_ + AdamW (train_nerf_regtr.py:96-102,232-239) over a fixed set of tensors. */
typedef struct drb_adamw drb_adamw;
int drb_adamw_create(int n, float* const* params_host, const long long* numels_host, drb_adamw** out);
void drb_adamw_destroy(drb_adamw* o);
/* max_grad_norm <= 0: no clipping.  No host synchronisation. */
int drb_adamw_step(drb_adamw* o, const float* const* grads_host, float lr, float beta1, float beta2,
                   float eps, float weight_decay, float max_grad_norm, drb_stream_t stream);
/* Total gradient norm seen by the last step (synchronous). */
int drb_adamw_grad_norm(drb_adamw* o, double* host_norm, drb_stream_t stream);
/* Copies exp_avg / exp_avg_sq of tensor i into caller storage (device, same numel); *step = steps taken. */
int drb_adamw_copy_state(drb_adamw* o, int i, float* exp_avg_out, float* exp_avg_sq_out, long long* step,
                         drb_stream_t stream);
int drb_adamw_set_step(drb_adamw* o, long long step);

/* ------------------------------------------------------------------------------------------
 * Whole-path engine: NeRFRegTr.forward (conerf/register/nerf_regtr.py:112-248) without Python in
 * the loop.  Parameters are addressed by the reference's state-dict key.
 * ---------------------------------------------------------------------------------------- */
typedef struct drb_engine drb_engine;

typedef struct drb_engine_config {
  int res_x, res_y, res_z;     /* grid resolution X, Y, Z                                      */
  int planes;                  /* 1 = bf16, 2 = split-bf16 (fp32-grade)                        */
  int num_downsample;          /* NeRFRegTr(num_downsample=6)                                  */
  float pos_emb_scaling;       /* NeRFRegTr(pos_emb_scaling=1.0)                               */
  int max_mask;                /* capacity for masked voxels per cloud                         */
  int training_bn;             /* 1: batch statistics + running-stat update (module.training)  */
} drb_engine_config;

int drb_engine_create(const drb_engine_config* cfg, drb_engine** out);
void drb_engine_destroy(drb_engine* e);
/* Parameter table (names are the reference's state_dict keys). */
int drb_engine_num_params(const drb_engine* e);
const char* drb_engine_param_name(const drb_engine* e, int i);
long long drb_engine_param_numel(const drb_engine* e, int i);
/* Binds fp32 device storage for parameter i (kept by reference: BN running buffers are updated in
 * place in training mode).  Call drb_engine_commit_params after all are bound or changed. */
int drb_engine_bind_param(drb_engine* e, int i, float* device_ptr);
int drb_engine_commit_params(drb_engine* e, drb_stream_t stream);
int drb_engine_set_training(drb_engine* e, int training_bn);
/* on (default): evaluate the level-1 FPN convolutions only on the output tiles the masked gather reads
 * (identical results at every voxel that is read); off: dense evaluation (debug taps of p1). */
int drb_engine_set_sparse_fpn(drb_engine* e, int on);
/* Several engines may share one set of bound parameters, one engine per CUDA stream (pipeline.py: pairs of a
 * batch in flight on different streams).  on (default): training-mode BatchNorm updates the bound running_mean /
 * running_var in place (nn.BatchNorm3d, momentum 0.1); off: this engine normalises with its batch statistics
 * but leaves the shared running buffers alone - exactly one engine of the group updates them, as rank 0's
 * buffers win among the replicas of a DistributedDataParallel job. */
int drb_engine_set_update_running(drb_engine* e, int on);

typedef struct drb_pair_io {
  const float* src_grid;   /* fp32 [1,7,Z,X,Y] view, element strides below                   */
  const float* tgt_grid;
  long long s_ch, s_z, s_x, s_y;          /* src strides                                     */
  long long t_ch, t_z, t_x, t_y;          /* tgt strides                                     */
  const long long* src_mask; int n_src_mask;
  const long long* tgt_mask; int n_tgt_mask;
} drb_pair_io;

/* Stage 1: FPN + gather + down-sampling.  Synchronises once to return the token counts. */
int drb_engine_encode(drb_engine* e, const drb_pair_io* io, int* host_n_src, int* host_n_tgt,
                      drb_stream_t stream);
/* Stage 2: transformer + decoder + Procrustes into caller buffers (sizes from stage 1):
 *   src_feats [6][ns][256], tgt_feats [6][nt][256], src_kp [ns][3], tgt_kp [nt][3],
 *   src_corr [6][ns][3], tgt_corr [6][nt][3], src_overlap [6][ns], tgt_overlap [6][nt],
 *   pose [6][3][4]. */
typedef struct drb_pair_out {
  float* src_feats; float* tgt_feats;
  float* src_kp; float* tgt_kp;
  float* src_corr; float* tgt_corr;
  float* src_overlap; float* tgt_overlap;
  float* pose;
} drb_pair_out;
int drb_engine_decode(drb_engine* e, const drb_pair_out* out, drb_stream_t stream);
/* Training: with grad mode on, encode / decode keep what the backward pass needs (per-layer convolution
 * outputs, BatchNorm statistics, transformer states) and drb_engine_backward accumulates the gradient of
 * every trainable parameter into the storage bound with drb_engine_bind_grad (zeroed by the caller).
 * One forward may be outstanding per engine. */
int drb_engine_set_grad_mode(drb_engine* e, int on);
int drb_engine_param_trainable(const drb_engine* e, int i);
int drb_engine_bind_grad(drb_engine* e, int i, float* device_ptr);
typedef struct drb_pair_grad {            /* gradients of the drb_pair_out tensors; NULL = zero */
  const float* d_src_feats; const float* d_tgt_feats;
  const float* d_src_corr; const float* d_tgt_corr;
  const float* d_src_overlap; const float* d_tgt_overlap;
  const float* d_pose;
} drb_pair_grad;
/* io / out: the arguments of the forward (same tensors, still alive). */
int drb_engine_backward(drb_engine* e, const drb_pair_io* io, const drb_pair_out* out,
                        const drb_pair_grad* grad, drb_stream_t stream);
/* 1: attention through drb_mha_tc_* (tcgen05), 0: the mma.sync kernel drb_mha_core. */
int drb_engine_set_tc_attention(drb_engine* e, int on);
/* Caps the down-sampler's stopping rule (grid_downsample.py:70,91; default 3000 tokens). */
int drb_engine_set_max_tokens(drb_engine* e, int max_total);

/* Debug / parity taps: copies a named intermediate ("c1".."c5", "p1".."p5", "rows") of grid
 * `which` (0 src, 1 tgt) as fp32 channels-last into dst; returns element count via *numel. */
int drb_engine_tap(drb_engine* e, const char* name, int which, float* dst, long long capacity,
                   long long* numel, drb_stream_t stream);
/* Kernel launches issued by this engine since creation (for bench.py's gpu_launches). */
long long drb_engine_launch_count(const drb_engine* e);
/* Roofline instrumentation: when on, every tcgen05 GEMM launch is bracketed by CUDA events on the
 * launching stream.  drb_engine_profile_read synchronises, returns the summed device time (ms), the
 * summed algorithmic FLOPs (2*M*N*K, dense) and the number of launches, and clears the records. */
int drb_engine_set_profile(drb_engine* e, int on);
int drb_engine_profile_read(drb_engine* e, double* igemm_ms, double* igemm_flops, long long* n);

#ifdef __cplusplus
}
#endif
#endif /* DREGB200_H_ */
