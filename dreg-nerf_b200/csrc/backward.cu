// Backward-pass kernels of the registration path (train_nerf_regtr.py:229 `loss.backward()` through
// NeRFRegTr.forward, conerf/register/nerf_regtr.py:112-248): everything that is not a tensor-core GEMM.
// The GEMM-shaped gradients (dgrad / wgrad of every Conv3d and Linear) live in igemm.cu / wgrad.cu.
//
//   BatchNorm3d backward (batch statistics or running statistics)      resnet3d.py:82-87,121
//   MaxPool3d(3,2,1) backward                                           resnet3d.py:123
//   nearest-upsample + add backward                                     feature_pyramid_net.py:58-61
//   trilinear-at-mask gather backward                                   nerf_regtr.py:138-147
//   col2im (strided convolutions lowered through im2col)                resnet3d.py:83,120,140-146
//   LayerNorm backward, attention soft-max backward, decoder soft-correspondence backward
//                                                                       transformer.py:225-299, nerf_regtr.py:292-306
//   overlap head backward                                               nerf_regtr.py:384-387
//   weighted Procrustes backward (closed form of the SVD derivative)    se3.py:89-140
//   fused clip_grad_norm_ + AdamW                                       train_nerf_regtr.py:96-102,232-239
#include "common.cuh"
#include "svd3.cuh"

#include <math.h>

#include <vector>

namespace drb {

static inline int grid_for(long long n, int block, int cap = 148 * 16) {
  long long g = (n + block - 1) / block;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// ------------------------------------------------------------------------------------------
// Gradient planes.  Gradients span many orders of magnitude, fp16 does not: in pair mode a tensor is
// multiplied by a power of two that maps its max |x| into [256, 512) before the hi/lo split, and the
// inverse is left in a device float for the consumer's epilogue (drb_conv3d_desc.acc_scale_dev,
// drb_wgrad_desc.scale_dev).  bf16 mode needs no scaling.
// slot: device {uint absmax bits, float inv_scale}.
// ------------------------------------------------------------------------------------------
__global__ void grad_absmax_kernel(const float* __restrict__ x, long long n, unsigned int* __restrict__ slot) {
  float m = 0.f;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long n4 = n >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = *(const float4*)(x + i * 4);
    m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
  }
  for (long long i = n4 * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    m = fmaxf(m, fabsf(x[i]));
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(slot, __float_as_uint(m));
}

__device__ __forceinline__ float pow2_scale_for(unsigned int bits) {
  const float mx = __uint_as_float(bits);
  if (!(mx > 0.f) || !isfinite(mx)) return 1.f;
  int e = 0;
  frexpf(mx, &e);                 // mx = f * 2^e, f in [0.5, 1)
  return ldexpf(1.f, 9 - e);      // mx * scale in [256, 512)
}

// rows x cols fp32 (pitch ld_in) -> planes (pitch ld_out >= cols, padding columns zeroed)
__global__ void grad_split_kernel(const float* __restrict__ x, long long rows, int cols, long long ld_in,
                                  long long ld_out, const unsigned int* __restrict__ slot_bits,
                                  float* __restrict__ inv_scale_out, plane_t* __restrict__ hi,
                                  plane_t* __restrict__ lo) {
  const bool pair = lo != nullptr;
  const float scale = (pair && slot_bits) ? pow2_scale_for(*slot_bits) : 1.f;
  if (inv_scale_out && blockIdx.x == 0 && threadIdx.x == 0) *inv_scale_out = 1.f / scale;
  const long long total = rows * ld_out;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long r = i / ld_out;
    const int c = (int)(i - r * ld_out);
    const float v = c < cols ? x[r * ld_in + c] * scale : 0.f;
    plane_t h, l;
    split16(v, pair, h, l);
    hi[i] = h;
    if (pair) lo[i] = l;
  }
}

// contiguous fast path: 4 values per thread
__global__ void grad_split4_kernel(const float* __restrict__ x, long long n4, const unsigned int* __restrict__ slot_bits,
                                   float* __restrict__ inv_scale_out, plane_t* __restrict__ hi, plane_t* __restrict__ lo) {
  const bool pair = lo != nullptr;
  const float scale = (pair && slot_bits) ? pow2_scale_for(*slot_bits) : 1.f;
  if (inv_scale_out && blockIdx.x == 0 && threadIdx.x == 0) *inv_scale_out = 1.f / scale;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = *(const float4*)(x + i * 4);
    plane_t h0, l0, h1, l1, h2, l2, h3, l3;
    split16(v.x * scale, pair, h0, l0); split16(v.y * scale, pair, h1, l1);
    split16(v.z * scale, pair, h2, l2); split16(v.w * scale, pair, h3, l3);
    *(uint2*)(hi + i * 4) = make_uint2(pack16x2(h0, h1), pack16x2(h2, h3));
    if (pair) *(uint2*)(lo + i * 4) = make_uint2(pack16x2(l0, l1), pack16x2(l2, l3));
  }
}

extern "C" int drb_grad_split(const float* x, long long rows, int cols, long long ld_in, long long ld_out,
                              void* hi, void* lo, float* slot, cudaStream_t stream) {
  DRB_REQUIRE(x && hi && slot && rows >= 0 && cols > 0 && ld_in >= cols && ld_out >= cols,
              "drb_grad_split: bad arguments");
  if (rows == 0) return 0;
  unsigned int* bits = (unsigned int*)slot;
  if (lo) {
    DRB_CUDA_OK(cudaMemsetAsync(bits, 0, sizeof(unsigned int), stream));
    // a pitched input is scanned over its whole extent: callers keep the padding columns at zero
    const long long span = (rows - 1) * ld_in + cols;
    grad_absmax_kernel<<<grid_for(span / 4 + 1, 256, 148 * 8), 256, 0, stream>>>(x, span, bits);
    DRB_LAUNCH_OK();
  }
  if (ld_in == cols && ld_out == cols && (rows * cols) % 4 == 0) {
    grad_split4_kernel<<<grid_for(rows * cols / 4, 256, 148 * 16), 256, 0, stream>>>(
        x, rows * cols / 4, lo ? bits : nullptr, slot + 1, (plane_t*)hi, (plane_t*)lo);
    DRB_LAUNCH_OK();
    return 0;
  }
  grad_split_kernel<<<grid_for(rows * ld_out, 256, 148 * 32), 256, 0, stream>>>(
      x, rows, cols, ld_in, ld_out, lo ? bits : nullptr, slot + 1, (plane_t*)hi, (plane_t*)lo);
  DRB_LAUNCH_OK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------
__global__ void add_inplace_kernel(float* __restrict__ dst, const float* __restrict__ src, long long n4) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 a = *(float4*)(dst + i * 4);
    const float4 b = *(const float4*)(src + i * 4);
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    *(float4*)(dst + i * 4) = a;
  }
}
extern "C" int drb_add_inplace(float* dst, const float* src, long long n, cudaStream_t stream) {
  DRB_REQUIRE(dst && src && n >= 0 && n % 4 == 0, "drb_add_inplace: bad arguments");
  if (n == 0) return 0;
  add_inplace_kernel<<<grid_for(n / 4, 256, 148 * 32), 256, 0, stream>>>(dst, src, n / 4);
  DRB_LAUNCH_OK();
  return 0;
}

// dy[i] = 0 where the saved ReLU output plane (hi) is zero (FFN hidden activations, transformer.py:291)
__global__ void relu_mask_plane_kernel(float* __restrict__ dy, const plane_t* __restrict__ hi, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    if ((hi[i] & 0x7fffu) == 0u) dy[i] = 0.f;
}
extern "C" int drb_relu_mask_plane(float* dy, const void* hi, long long n, cudaStream_t stream) {
  DRB_REQUIRE(dy && hi && n >= 0, "drb_relu_mask_plane: bad arguments");
  if (n == 0) return 0;
  relu_mask_plane_kernel<<<grid_for(n, 256, 148 * 32), 256, 0, stream>>>(dy, (const plane_t*)hi, n);
  DRB_LAUNCH_OK();
  return 0;
}

// out[c] += sum over rows of x[r][c]  (bias gradients).  Block = 16 channel groups (float4) x 16 row lanes.
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ x, long long rows, int c, long long ld,
                                                      int rows_per_block, float* __restrict__ out) {
  const int cg = threadIdx.x & 15, rl = threadIdx.x >> 4;
  const int ch = blockIdx.y * 64 + cg * 4;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > rows) r1 = rows;
  float s[4] = {0.f, 0.f, 0.f, 0.f};
  if (ch + 3 < c) {
#pragma unroll 4
    for (long long r = r0 + rl; r < r1; r += 16) {
      const float4 v = *(const float4*)(x + r * ld + ch);
      s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
    }
  } else {
    for (int j = 0; j < 4; ++j)
      if (ch + j < c)
        for (long long r = r0 + rl; r < r1; r += 16) s[j] += x[r * ld + ch + j];
  }
  __shared__ float sh[16][64];
#pragma unroll
  for (int j = 0; j < 4; ++j) sh[rl][cg * 4 + j] = s[j];
  __syncthreads();
  if (threadIdx.x < 64) {
    const int cc = blockIdx.y * 64 + threadIdx.x;
    if (cc < c) {
      double t = 0.0;
#pragma unroll
      for (int k = 0; k < 16; ++k) t += (double)sh[k][threadIdx.x];
      atomicAdd(out + cc, (float)t);
    }
  }
}
extern "C" int drb_colsum_add(const float* x, long long rows, int c, long long ld, float* out,
                              cudaStream_t stream) {
  DRB_REQUIRE(x && out && c > 0 && ld >= c && ld % 4 == 0, "drb_colsum_add: bad arguments");
  if (rows == 0) return 0;
  int rpb = 256;
  while ((rows + rpb - 1) / rpb > 2048) rpb *= 2;
  dim3 grid((unsigned)((rows + rpb - 1) / rpb), (unsigned)((c + 63) / 64));
  colsum_kernel<<<grid, 256, 0, stream>>>(x, rows, c, ld, rpb, out);
  DRB_LAUNCH_OK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// BatchNorm3d backward.  raw = the convolution output the forward normalised, [g][m][c].
//   pass 1 (reduce): S1 = sum dy, S2 = sum dy * xhat per (g, c) over the ReLU-masked dy (mask: post > 0 when the
//           forward kept the activation in fp32, else sign of fma(raw, scale, shift) - the forward's own
//           expression); read only.
//   pass 2 (apply):  batch statistics:  dx = gamma rstd (dy - S1/m - xhat S2/m)
//                    running statistics: dx = gamma rstd dy
//           dgamma += sum_g S2, dbeta += sum_g S1.
// ------------------------------------------------------------------------------------------
// Block = 16 channel groups (float4 = 4 channels each: 64 channels) x 16 row lanes.  The reduce pass only reads
// (the mask is recomputed by the apply pass): per-thread fp32 partials over <= 64 rows, fp64 across the block.
__device__ __forceinline__ float4 bn_mask4(float4 d, float4 x, const float* __restrict__ post4, float4 sc, float4 sf,
                                           int relu) {
  if (!relu) return d;
  if (post4) {
    const float4 p = *(const float4*)post4;
    d.x = p.x > 0.f ? d.x : 0.f; d.y = p.y > 0.f ? d.y : 0.f; d.z = p.z > 0.f ? d.z : 0.f; d.w = p.w > 0.f ? d.w : 0.f;
  } else {
    d.x = fmaf(x.x, sc.x, sf.x) > 0.f ? d.x : 0.f; d.y = fmaf(x.y, sc.y, sf.y) > 0.f ? d.y : 0.f;
    d.z = fmaf(x.z, sc.z, sf.z) > 0.f ? d.z : 0.f; d.w = fmaf(x.w, sc.w, sf.w) > 0.f ? d.w : 0.f;
  }
  return d;
}

__global__ void __launch_bounds__(256)
bn_bwd_reduce_kernel(const float* __restrict__ dy, const float* __restrict__ raw, const float* __restrict__ post,
                     const float* __restrict__ scale, const float* __restrict__ shift,
                     const float* __restrict__ mean, const float* __restrict__ rstd, int relu, long long m, int c,
                     int rows_per_block, double* __restrict__ sums) {
  const int g = blockIdx.z;
  const int cg = threadIdx.x & 15, rl = threadIdx.x >> 4;
  const int ch = blockIdx.y * 64 + cg * 4;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > m) r1 = m;
  float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
  if (ch < c) {
    const long long gc = (long long)g * c + ch;
    const float4 mu = *(const float4*)(mean + gc), rs = *(const float4*)(rstd + gc);
    float4 sc = make_float4(0.f, 0.f, 0.f, 0.f), sf = sc;
    if (scale) { sc = *(const float4*)(scale + gc); sf = *(const float4*)(shift + gc); }
    const long long base = ((long long)g * m) * c + ch;
#pragma unroll 4
    for (long long r = r0 + rl; r < r1; r += 16) {
      const long long i = base + r * c;
      const float4 x = *(const float4*)(raw + i);
      const float4 d = bn_mask4(*(const float4*)(dy + i), x, post ? post + i : nullptr, sc, sf, relu);
      s1[0] += d.x; s1[1] += d.y; s1[2] += d.z; s1[3] += d.w;
      s2[0] = fmaf(d.x, (x.x - mu.x) * rs.x, s2[0]); s2[1] = fmaf(d.y, (x.y - mu.y) * rs.y, s2[1]);
      s2[2] = fmaf(d.z, (x.z - mu.z) * rs.z, s2[2]); s2[3] = fmaf(d.w, (x.w - mu.w) * rs.w, s2[3]);
    }
  }
  __shared__ float sh[2][16][64];
#pragma unroll
  for (int j = 0; j < 4; ++j) { sh[0][rl][cg * 4 + j] = s1[j]; sh[1][rl][cg * 4 + j] = s2[j]; }
  __syncthreads();
  if (threadIdx.x < 128) {
    const int which = threadIdx.x >> 6, cl = threadIdx.x & 63;
    const int cc = blockIdx.y * 64 + cl;
    if (cc < c) {
      double t = 0.0;
#pragma unroll
      for (int k = 0; k < 16; ++k) t += (double)sh[which][k][cl];
      atomicAdd(&sums[((long long)g * c + cc) * 2 + which], t);
    }
  }
}

// dx = A dy_masked + B raw + C; optionally writes the masked dy back (the residual branch's gradient).
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const float* dy, const float* __restrict__ raw, const float* __restrict__ post,
                    const float* __restrict__ scale, const float* __restrict__ shift, const double* __restrict__ sums,
                    const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ rstd,
                    int relu, int training, int g_total, long long m, int c, int rows_per_block, float* dy_masked_out,
                    float* dx, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int g = blockIdx.z;
  const int cg = threadIdx.x & 15, rl = threadIdx.x >> 4;
  const int ch = blockIdx.y * 64 + cg * 4;
  if (ch >= c) return;
  const long long gc = (long long)g * c + ch;
  float A[4], B[4], C[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float ga = gamma ? gamma[ch + j] : 1.f;
    const float mu = mean[gc + j], rs = rstd[gc + j];
    const double S1 = sums[(gc + j) * 2], S2 = sums[(gc + j) * 2 + 1];
    A[j] = ga * rs; B[j] = 0.f; C[j] = 0.f;
    if (training) {
      const double inv_m = 1.0 / (double)m;
      B[j] = (float)(-(double)ga * (double)rs * (double)rs * S2 * inv_m);
      C[j] = (float)(-(double)ga * (double)rs * S1 * inv_m + (double)ga * (double)rs * (double)rs * (double)mu * S2 * inv_m);
    }
    if (blockIdx.x == 0 && g == 0 && rl == 0) {
      double t1 = 0.0, t2 = 0.0;
      for (int gi = 0; gi < g_total; ++gi) {
        t1 += sums[((long long)gi * c + ch + j) * 2];
        t2 += sums[((long long)gi * c + ch + j) * 2 + 1];
      }
      if (dbeta) dbeta[ch + j] += (float)t1;
      if (dgamma) dgamma[ch + j] += (float)t2;
    }
  }
  float4 sc = make_float4(0.f, 0.f, 0.f, 0.f), sf = sc;
  if (scale) { sc = *(const float4*)(scale + gc); sf = *(const float4*)(shift + gc); }
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > m) r1 = m;
  const long long base = ((long long)g * m) * c + ch;
#pragma unroll 4
  for (long long r = r0 + rl; r < r1; r += 16) {
    const long long i = base + r * c;
    const float4 x = *(const float4*)(raw + i);
    const float4 d = bn_mask4(*(const float4*)(dy + i), x, post ? post + i : nullptr, sc, sf, relu);
    if (dy_masked_out) *(float4*)(dy_masked_out + i) = d;
    float4 o;
    o.x = fmaf(A[0], d.x, fmaf(B[0], x.x, C[0])); o.y = fmaf(A[1], d.y, fmaf(B[1], x.y, C[1]));
    o.z = fmaf(A[2], d.z, fmaf(B[2], x.z, C[2])); o.w = fmaf(A[3], d.w, fmaf(B[3], x.w, C[3]));
    *(float4*)(dx + i) = o;
  }
}

// dy is masked in place when relu != 0 and dx != dy (the masked gradient is what a residual branch receives);
// dx may alias dy.  sums: [g][c][2] doubles of scratch.  c must be a multiple of 4.
extern "C" int drb_bn_backward(float* dy, const float* raw, const float* post, const float* scale,
                               const float* shift, const float* mean, const float* rstd, const float* gamma,
                               int relu, int training, int g, long long m, int c, double* sums, float* dx,
                               float* dgamma, float* dbeta, cudaStream_t stream) {
  DRB_REQUIRE(dy && raw && mean && rstd && sums && dx && g > 0 && m > 0 && c > 0 && c % 4 == 0,
              "drb_bn_backward: bad arguments");
  DRB_REQUIRE(!relu || post || (scale && shift), "drb_bn_backward: ReLU mask needs post or scale/shift");
  DRB_CUDA_OK(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * (size_t)g * c, stream));
  int rpb = 256;
  while ((m + rpb - 1) / rpb > 2048) rpb *= 2;
  dim3 grid((unsigned)((m + rpb - 1) / rpb), (unsigned)((c + 63) / 64), (unsigned)g);
  bn_bwd_reduce_kernel<<<grid, 256, 0, stream>>>(dy, raw, post, scale, shift, mean, rstd, relu, m, c, rpb, sums);
  DRB_LAUNCH_OK();
  float* masked_out = (relu && dx != dy) ? dy : nullptr;
  bn_bwd_apply_kernel<<<grid, 256, 0, stream>>>(dy, raw, post, scale, shift, sums, gamma, mean, rstd, relu, training, g,
                                                m, c, rpb, masked_out, dx, dgamma, dbeta);
  DRB_LAUNCH_OK();
  return 0;
}

// Forward-side companion: mean / rstd / scale / shift per (g, c) exactly as bn_apply_kernel derives them.
__global__ void bn_save_kernel(const double* __restrict__ accum, int g, long long m, int c,
                               const float* __restrict__ gamma, const float* __restrict__ beta,
                               const float* __restrict__ running_mean, const float* __restrict__ running_var,
                               int training, float eps, float* __restrict__ mean, float* __restrict__ rstd,
                               float* __restrict__ scale, float* __restrict__ shift) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g * c) return;
  const int ch = i % c;
  const float ga = gamma ? gamma[ch] : 1.f, be = beta ? beta[ch] : 0.f;
  float mu, rs;
  if (training) {
    const double mean_d = accum[(long long)i * 2] / (double)m;
    double var = accum[(long long)i * 2 + 1] / (double)m - mean_d * mean_d;
    if (var < 0.0) var = 0.0;
    rs = (float)(1.0 / sqrt(var + (double)eps));
    mu = (float)mean_d;
  } else {
    rs = 1.f / sqrtf(running_var[ch] + eps);
    mu = running_mean[ch];
  }
  const float sc = ga * rs;
  mean[i] = mu; rstd[i] = rs;
  scale[i] = sc; shift[i] = be - mu * sc;
}
// Must run BEFORE drb_bn_apply in training mode when running statistics matter?  No: it only reads accum
// (training) or the running buffers (eval, which drb_bn_apply does not modify).
extern "C" int drb_bn_save_stats(const double* accum, int g, long long m, int c, const float* gamma,
                                 const float* beta, const float* running_mean, const float* running_var,
                                 int training, float eps, float* mean, float* rstd, float* scale, float* shift,
                                 cudaStream_t stream) {
  DRB_REQUIRE(mean && rstd && scale && shift && g > 0 && c > 0, "drb_bn_save_stats: bad arguments");
  DRB_REQUIRE(training ? accum != nullptr : (running_mean && running_var), "drb_bn_save_stats: missing statistics");
  bn_save_kernel<<<cdiv((long long)g * c, 128), 128, 0, stream>>>(accum, g, m, c, gamma, beta, running_mean,
                                                               running_var, training, eps, mean, rstd, scale, shift);
  DRB_LAUNCH_OK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// MaxPool3d(3, 2, 1) backward: the gradient of an output goes to the first maximum of its window in
// (z, y, x) scan order (ATen max_pool3d_with_indices: `val > maxval`).  dx must be zeroed by the caller.
// ------------------------------------------------------------------------------------------
__global__ void maxpool_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dout, int g, int d,
                                   int h, int w, int c, int od, int oh, int ow, float* __restrict__ dx) {
  const long long total = (long long)g * od * oh * ow * c;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const float gout = dout[i];
    if (gout == 0.f) continue;
    const int ch = (int)(i % c);
    long long r = i / c;
    const int ox = (int)(r % ow); r /= ow;
    const int oy = (int)(r % oh); r /= oh;
    const int oz = (int)(r % od); r /= od;
    const int gi = (int)r;
    float best = -INFINITY;
    long long arg = -1;
    for (int kz = 0; kz < 3; ++kz) {
      const int iz = oz * 2 - 1 + kz;
      if (iz < 0 || iz >= d) continue;
      for (int ky = 0; ky < 3; ++ky) {
        const int iy = oy * 2 - 1 + ky;
        if (iy < 0 || iy >= h) continue;
        for (int kx = 0; kx < 3; ++kx) {
          const int ix = ox * 2 - 1 + kx;
          if (ix < 0 || ix >= w) continue;
          const long long idx = ((((long long)gi * d + iz) * h + iy) * w + ix) * c + ch;
          const float v = x[idx];
          if (v > best || arg < 0) { best = v; arg = idx; }
        }
      }
    }
    if (arg >= 0) atomicAdd(dx + arg, gout);
  }
}
extern "C" int drb_maxpool3d_backward(const float* x, const float* dout, int g, int d, int h, int w, int c,
                                      float* dx, cudaStream_t stream) {
  DRB_REQUIRE(x && dout && dx && g > 0 && d > 0 && h > 0 && w > 0 && c > 0, "drb_maxpool3d_backward: bad arguments");
  const int od = (d - 1) / 2 + 1, oh = (h - 1) / 2 + 1, ow = (w - 1) / 2 + 1;
  DRB_CUDA_OK(cudaMemsetAsync(dx, 0, sizeof(float) * (size_t)g * d * h * w * c, stream));
  const long long total = (long long)g * od * oh * ow * c;
  maxpool_bwd_kernel<<<grid_for(total, 256, 148 * 32), 256, 0, stream>>>(x, dout, g, d, h, w, c, od, oh, ow, dx);
  DRB_LAUNCH_OK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// FeaturePyramid_v1._upsample + add backward: dcoarse[z][y][x] = sum of the (cropped) 2^3 children of
// dsum; dlateral == dsum (no kernel).
// ------------------------------------------------------------------------------------------
__global__ void upsample_add_bwd_kernel(const float* __restrict__ dsum, int g, int d, int h, int w, int c,
                                        int dc, int hc, int wc, float* __restrict__ dcoarse) {
  const int c4 = c >> 2;
  const long long total4 = (long long)g * dc * hc * wc * c4;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += stride) {
    const int cc = (int)(i % c4) * 4;
    long long r = i / c4;
    const int x = (int)(r % wc); r /= wc;
    const int y = (int)(r % hc); r /= hc;
    const int z = (int)(r % dc); r /= dc;
    const int gi = (int)r;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int a = 0; a < 2; ++a) {
      const int zz = 2 * z + a;
      if (zz >= d) continue;
      for (int b = 0; b < 2; ++b) {
        const int yy = 2 * y + b;
        if (yy >= h) continue;
        for (int e = 0; e < 2; ++e) {
          const int xx = 2 * x + e;
          if (xx >= w) continue;
          const float4 v = *(const float4*)(dsum + ((((long long)gi * d + zz) * h + yy) * w + xx) * c + cc);
          acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
      }
    }
    *(float4*)(dcoarse + i * 4) = acc;
  }
}
extern "C" int drb_upsample2_add_backward(const float* dsum, int g, int d, int h, int w, int c, int dc, int hc,
                                          int wc, float* dcoarse, cudaStream_t stream) {
  DRB_REQUIRE(dsum && dcoarse && c % 4 == 0 && d <= 2 * dc && h <= 2 * hc && w <= 2 * wc,
              "drb_upsample2_add_backward: bad arguments");
  const long long total4 = (long long)g * dc * hc * wc * (c / 4);
  upsample_add_bwd_kernel<<<grid_for(total4, 256, 148 * 32), 256, 0, stream>>>(dsum, g, d, h, w, c, dc, hc, wc,
                                                                             dcoarse);
  DRB_LAUNCH_OK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// Trilinear-at-mask gather backward: scatter-add of the 8 corner weights (same index arithmetic as
// trilinear_gather_kernel).  drows: [k][ld] with the feature gradient at columns [col0, col0 + c).
// dp1 (one grid, [dc][hc][wc][c]) must be zeroed by the caller.
// ------------------------------------------------------------------------------------------
__global__ void trilinear_scatter_kernel(const float* __restrict__ drows, int ld, int col0, int dc, int hc,
                                         int wc, int c, int X, int Y, int Z, const long long* __restrict__ mask,
                                         int k, float* __restrict__ dp1) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= k) return;
  const long long idx = mask[warp];
  const int z = (int)(idx % Z);
  const int y = (int)((idx / Z) % Y);
  const int x = (int)(idx / ((long long)Z * Y));
  const float sd = Z > 1 ? (float)(dc - 1) / (float)(Z - 1) : 0.f;
  const float sh = X > 1 ? (float)(hc - 1) / (float)(X - 1) : 0.f;
  const float sw = Y > 1 ? (float)(wc - 1) / (float)(Y - 1) : 0.f;
  const float fd = sd * (float)z, fh = sh * (float)x, fw = sw * (float)y;
  const int d0 = (int)fd, h0 = (int)fh, w0 = (int)fw;
  const int d1 = d0 + (d0 < dc - 1 ? 1 : 0), h1 = h0 + (h0 < hc - 1 ? 1 : 0), w1 = w0 + (w0 < wc - 1 ? 1 : 0);
  const float ld1 = fd - (float)d0, ld0 = 1.f - ld1;
  const float lh1 = fh - (float)h0, lh0 = 1.f - lh1;
  const float lw1 = fw - (float)w0, lw0 = 1.f - lw1;
  const int ds[2] = {d0, d1}, hs[2] = {h0, h1}, ws[2] = {w0, w1};
  const float wd[2] = {ld0, ld1}, wh[2] = {lh0, lh1}, ww[2] = {lw0, lw1};
  const float* src = drows + (long long)warp * ld + col0;
  for (int cc = lane * 4; cc < c; cc += 128) {
    const float4 gv = *(const float4*)(src + cc);
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const float wt = wd[a] * wh[b] * ww[e];
          if (wt == 0.f) continue;
          float* dst = dp1 + (((long long)ds[a] * hc + hs[b]) * wc + ws[e]) * c + cc;
          atomicAdd(dst + 0, wt * gv.x); atomicAdd(dst + 1, wt * gv.y);
          atomicAdd(dst + 2, wt * gv.z); atomicAdd(dst + 3, wt * gv.w);
        }
  }
}
extern "C" int drb_trilinear_gather_backward(const float* drows, int ld, int col0, int dc, int hc, int wc, int c,
                                             int X, int Y, int Z, const long long* mask, int k, float* dp1,
                                             cudaStream_t stream) {
  DRB_REQUIRE(drows && mask && dp1 && c % 4 == 0 && ld % 4 == 0 && col0 % 4 == 0 && ld >= col0 + c,
              "drb_trilinear_gather_backward: bad arguments");
  if (k == 0) return 0;
  trilinear_scatter_kernel<<<cdiv(k, 8), 256, 0, stream>>>(drows, ld, col0, dc, hc, wc, c, X, Y, Z, mask, k, dp1);
  DRB_LAUNCH_OK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// col2im: dx[g][iz][iy][ix][ch] = sum over taps of dcol[(g, oz, oy, ox)][tap * c + ch] (+ residual),
// the adjoint of im2col_kernel (gather form: no atomics).
// ------------------------------------------------------------------------------------------
__global__ void col2im_kernel(const float* __restrict__ dcol, int g, int c, int d, int h, int w, int k, int stride,
                              int pad, int kpad, int od, int oh, int ow, const float* __restrict__ residual,
                              float* __restrict__ dx) {
  const long long total = (long long)g * d * h * w * c;
  const long long gstride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gstride) {
    const int ch = (int)(i % c);
    long long r = i / c;
    const int ix = (int)(r % w); r /= w;
    const int iy = (int)(r % h); r /= h;
    const int iz = (int)(r % d); r /= d;
    const int gi = (int)r;
    float acc = residual ? residual[i] : 0.f;
    for (int kz = 0; kz < k; ++kz) {
      const int tz = iz + pad - kz;
      if (tz < 0 || tz % stride) continue;
      const int oz = tz / stride;
      if (oz >= od) continue;
      for (int ky = 0; ky < k; ++ky) {
        const int ty = iy + pad - ky;
        if (ty < 0 || ty % stride) continue;
        const int oy = ty / stride;
        if (oy >= oh) continue;
        for (int kx = 0; kx < k; ++kx) {
          const int tx = ix + pad - kx;
          if (tx < 0 || tx % stride) continue;
          const int ox = tx / stride;
          if (ox >= ow) continue;
          const long long row = (((long long)gi * od + oz) * oh + oy) * ow + ox;
          acc += dcol[row * kpad + ((kz * k + ky) * k + kx) * c + ch];
        }
      }
    }
    dx[i] = acc;
  }
}
// four channels per thread (16-byte loads / stores, a quarter of the index arithmetic); same summation order per channel
__global__ void __launch_bounds__(256) col2im4_kernel(const float* __restrict__ dcol, int g, int c, int d, int h, int w,
                                                      int k, int stride, int pad, int kpad, int od, int oh, int ow,
                                                      const float* __restrict__ residual, float* __restrict__ dx) {
  const int c4n = c >> 2;
  const long long total = (long long)g * d * h * w * c4n;
  const long long gstride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gstride) {
    const int ch = (int)(i % c4n) * 4;
    int r = (int)(i / c4n);                       // voxels < 2^31 (checked by the host)
    const int ix = r % w; r /= w;
    const int iy = r % h; r /= h;
    const int iz = r % d; r /= d;
    const int gi = r;
    float4 acc = residual ? *(const float4*)(residual + i * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int kz = 0; kz < k; ++kz) {
      const int tz = iz + pad - kz;
      if (tz < 0 || tz % stride) continue;
      const int oz = tz / stride;
      if (oz >= od) continue;
      for (int ky = 0; ky < k; ++ky) {
        const int ty = iy + pad - ky;
        if (ty < 0 || ty % stride) continue;
        const int oy = ty / stride;
        if (oy >= oh) continue;
        for (int kx = 0; kx < k; ++kx) {
          const int tx = ix + pad - kx;
          if (tx < 0 || tx % stride) continue;
          const int ox = tx / stride;
          if (ox >= ow) continue;
          const long long row = (((long long)gi * od + oz) * oh + oy) * ow + ox;
          const float4 v = *(const float4*)(dcol + row * kpad + ((kz * k + ky) * k + kx) * c + ch);
          acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
      }
    }
    *(float4*)(dx + i * 4) = acc;
  }
}
extern "C" int drb_col2im(const float* dcol, int g, int c, int d, int h, int w, int k, int stride, int pad,
                          int kpad, const float* residual, float* dx, cudaStream_t stream) {
  DRB_REQUIRE(dcol && dx && g > 0 && c > 0 && k >= 1 && stride >= 1 && kpad >= k * k * k * c, "drb_col2im: bad arguments");
  const int od = (d + 2 * pad - k) / stride + 1, oh = (h + 2 * pad - k) / stride + 1, ow = (w + 2 * pad - k) / stride + 1;
  const long long total = (long long)g * d * h * w * c;
  const bool al = ((uintptr_t)dcol & 15) == 0 && ((uintptr_t)dx & 15) == 0 && (!residual || ((uintptr_t)residual & 15) == 0);
  if (c % 4 == 0 && kpad % 4 == 0 && al && (long long)g * d * h * w < (1LL << 31)) {
    col2im4_kernel<<<grid_for(total / 4, 256, 148 * 32), 256, 0, stream>>>(dcol, g, c, d, h, w, k, stride, pad, kpad, od,
                                                                        oh, ow, residual, dx);
    DRB_LAUNCH_OK();
    return 0;
  }
  col2im_kernel<<<grid_for(total, 256, 148 * 32), 256, 0, stream>>>(dcol, g, c, d, h, w, k, stride, pad, kpad, od, oh,
                                                                   ow, residual, dx);
  DRB_LAUNCH_OK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// LayerNorm(256) backward, one warp per row.  dx_accum[row] += dLN/dx (the residual stream's gradient is
// accumulated in place); dgamma / dbeta += column sums through shared memory + atomics.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) layernorm256_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                                int n, const float* __restrict__ gamma,
                                                                float* __restrict__ dx_accum, int overwrite,
                                                                float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int warp_in_block = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = lane * 4, c1 = 128 + lane * 4;
  const float4 g0 = *(const float4*)(gamma + c0), g1 = *(const float4*)(gamma + c1);
  const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
  float pg[8] = {0, 0, 0, 0, 0, 0, 0, 0}, pb[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int row = blockIdx.x * 8 + warp_in_block; row < n; row += gridDim.x * 8) {
    const float* xr = x + (long long)row * 256;
    const float* dr = dy + (long long)row * 256;
    const float4 a = *(const float4*)(xr + c0), b = *(const float4*)(xr + c1);
    const float4 da = *(const float4*)(dr + c0), db = *(const float4*)(dr + c1);
    float s = a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w;
    s = warp_sum(s);
    const float mean = s * (1.f / 256.f);
    float v[8] = {a.x - mean, a.y - mean, a.z - mean, a.w - mean, b.x - mean, b.y - mean, b.z - mean, b.w - mean};
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) ss += v[i] * v[i];
    ss = warp_sum(ss);
    const float rstd = rsqrtf(ss * (1.f / 256.f) + 1e-5f);
    const float d[8] = {da.x, da.y, da.z, da.w, db.x, db.y, db.z, db.w};
    float dxh[8], m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      v[i] *= rstd;                       // xhat
      dxh[i] = d[i] * gg[i];
      m1 += dxh[i];
      m2 += dxh[i] * v[i];
      pg[i] += d[i] * v[i];
      pb[i] += d[i];
    }
    m1 = warp_sum(m1) * (1.f / 256.f);
    m2 = warp_sum(m2) * (1.f / 256.f);
    float o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = rstd * (dxh[i] - m1 - v[i] * m2);
    float* out = dx_accum + (long long)row * 256;
    if (!overwrite) {
      const float4 p0 = *(const float4*)(out + c0), p1 = *(const float4*)(out + c1);
      o[0] += p0.x; o[1] += p0.y; o[2] += p0.z; o[3] += p0.w;
      o[4] += p1.x; o[5] += p1.y; o[6] += p1.z; o[7] += p1.w;
    }
    *(float4*)(out + c0) = make_float4(o[0], o[1], o[2], o[3]);
    *(float4*)(out + c1) = make_float4(o[4], o[5], o[6], o[7]);
  }
  __shared__ float shg[8][256], shb[8][256];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    shg[warp_in_block][c0 + i] = pg[i]; shg[warp_in_block][c1 + i] = pg[4 + i];
    shb[warp_in_block][c0 + i] = pb[i]; shb[warp_in_block][c1 + i] = pb[4 + i];
  }
  __syncthreads();
  const int ch = threadIdx.x;
  float tg = 0.f, tb = 0.f;
#pragma unroll
  for (int wv = 0; wv < 8; ++wv) { tg += shg[wv][ch]; tb += shb[wv][ch]; }
  if (dgamma) atomicAdd(dgamma + ch, tg);
  if (dbeta) atomicAdd(dbeta + ch, tb);
}
extern "C" int drb_layernorm256_backward(const float* x, const float* dy, int n, const float* gamma, float* dx,
                                         int overwrite, float* dgamma, float* dbeta, cudaStream_t stream) {
  DRB_REQUIRE(x && dy && gamma && dx, "drb_layernorm256_backward: bad arguments");
  if (n == 0) return 0;
  int grid = cdiv(n, 8);
  if (grid > 148 * 2) grid = 148 * 2;
  layernorm256_bwd_kernel<<<grid, 256, 0, stream>>>(x, dy, n, gamma, dx, overwrite, dgamma, dbeta);
  DRB_LAUNCH_OK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// Overlap head backward: o = sigmoid(w . f + b).  dfeat[row] += dlogit * w; dw += sum dlogit * f; db += sum.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) overlap_bwd_kernel(const float* __restrict__ feat, const float* __restrict__ ov,
                                                           const float* __restrict__ dov, int n,
                                                           const float* __restrict__ w, float* __restrict__ dfeat,
                                                           float* __restrict__ dw, float* __restrict__ db) {
  const int warp_in_block = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = lane * 4, c1 = 128 + lane * 4;
  const float4 w0 = *(const float4*)(w + c0), w1 = *(const float4*)(w + c1);
  float pw[8] = {0, 0, 0, 0, 0, 0, 0, 0}, pb = 0.f;
  for (int row = blockIdx.x * 8 + warp_in_block; row < n; row += gridDim.x * 8) {
    const float o = ov[row];
    const float dl = dov[row] * o * (1.f - o);
    const float* f = feat + (long long)row * 256;
    const float4 a = *(const float4*)(f + c0), b = *(const float4*)(f + c1);
    pw[0] += dl * a.x; pw[1] += dl * a.y; pw[2] += dl * a.z; pw[3] += dl * a.w;
    pw[4] += dl * b.x; pw[5] += dl * b.y; pw[6] += dl * b.z; pw[7] += dl * b.w;
    if (lane == 0) pb += dl;
    float* df = dfeat + (long long)row * 256;
    float4 p0 = *(float4*)(df + c0), p1 = *(float4*)(df + c1);
    p0.x += dl * w0.x; p0.y += dl * w0.y; p0.z += dl * w0.z; p0.w += dl * w0.w;
    p1.x += dl * w1.x; p1.y += dl * w1.y; p1.z += dl * w1.z; p1.w += dl * w1.w;
    *(float4*)(df + c0) = p0;
    *(float4*)(df + c1) = p1;
  }
  __shared__ float shw[8][256];
  __shared__ float shb[8];
#pragma unroll
  for (int i = 0; i < 4; ++i) { shw[warp_in_block][c0 + i] = pw[i]; shw[warp_in_block][c1 + i] = pw[4 + i]; }
  if (lane == 0) shb[warp_in_block] = pb;
  __syncthreads();
  const int ch = threadIdx.x;
  float t = 0.f;
#pragma unroll
  for (int wv = 0; wv < 8; ++wv) t += shw[wv][ch];
  atomicAdd(dw + ch, t);
  if (ch == 0) {
    float tb = 0.f;
    for (int wv = 0; wv < 8; ++wv) tb += shb[wv];
    atomicAdd(db, tb);
  }
}
extern "C" int drb_overlap_sigmoid_backward(const float* feat, const float* ov, const float* dov, int n,
                                            const float* w, float* dfeat, float* dw, float* db,
                                            cudaStream_t stream) {
  DRB_REQUIRE(feat && ov && dov && w && dfeat && dw && db, "drb_overlap_sigmoid_backward: null argument");
  if (n == 0) return 0;
  int grid = cdiv(n, 8);
  if (grid > 148 * 2) grid = 148 * 2;
  overlap_bwd_kernel<<<grid, 256, 0, stream>>>(feat, ov, dov, n, w, dfeat, dw, db);
  DRB_LAUNCH_OK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// Row soft-max helpers for the attention backward (one warp per row of a [batch][rows][ld] matrix).
//   softmax_rows:      s <- softmax(s * scale) over the first nk columns (padding columns <- 0)
//   softmax_bwd_rows:  dp <- p * (dp - sum_j p_j dp_j) * scale
//   corr_bwd_rows:     s <- p * (g - sum_j p_j g_j) with g_j = dcorr_row . xyz_j   (decoder, values = xyz)
// ------------------------------------------------------------------------------------------
__global__ void softmax_rows_kernel(float* __restrict__ s, long long rows, int nk, int ld, float scale) {
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float* sr = s + row * ld;
  float mx = -INFINITY;
  for (int j = lane; j < nk; j += 32) mx = fmaxf(mx, sr[j] * scale);
  mx = warp_max(mx);
  float l = 0.f;
  for (int j = lane; j < nk; j += 32) {
    const float p = expf(sr[j] * scale - mx);
    sr[j] = p;
    l += p;
  }
  l = warp_sum(l);
  const float inv = 1.f / l;
  for (int j = lane; j < ld; j += 32) sr[j] = j < nk ? sr[j] * inv : 0.f;
}
__global__ void softmax_bwd_rows_kernel(const float* __restrict__ p, float* __restrict__ dp, long long rows, int nk,
                                        int ld, float scale) {
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* pr = p + row * ld;
  float* dr = dp + row * ld;
  float acc = 0.f;
  for (int j = lane; j < nk; j += 32) acc = fmaf(pr[j], dr[j], acc);
  acc = warp_sum(acc);
  for (int j = lane; j < ld; j += 32) dr[j] = j < nk ? pr[j] * (dr[j] - acc) * scale : 0.f;
}
__global__ void corr_bwd_rows_kernel(float* __restrict__ s, int nq, int nk, int ld, const float* __restrict__ xyz,
                                     int ld_xyz, const float* __restrict__ dcorr) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= nq) return;
  float* sr = s + (long long)row * ld;
  const float gx = dcorr[(long long)row * 3], gy = dcorr[(long long)row * 3 + 1], gz = dcorr[(long long)row * 3 + 2];
  float mx = -INFINITY;
  for (int j = lane; j < nk; j += 32) mx = fmaxf(mx, sr[j]);
  mx = warp_max(mx);
  float l = 0.f, acc = 0.f;
  for (int j = lane; j < nk; j += 32) {
    const float p = expf(sr[j] - mx);
    sr[j] = p;
    l += p;
    const float* pj = xyz + (long long)j * ld_xyz;
    acc = fmaf(p, gx * pj[0] + gy * pj[1] + gz * pj[2], acc);
  }
  l = warp_sum(l);
  acc = warp_sum(acc);
  const float inv = 1.f / l;
  const float cbar = acc * inv;
  for (int j = lane; j < ld; j += 32) {
    float v = 0.f;
    if (j < nk) {
      const float* pj = xyz + (long long)j * ld_xyz;
      v = sr[j] * inv * (gx * pj[0] + gy * pj[1] + gz * pj[2] - cbar);
    }
    sr[j] = v;
  }
}
extern "C" int drb_softmax_rows(float* s, long long rows, int nk, int ld, float scale, cudaStream_t stream) {
  DRB_REQUIRE(s && nk > 0 && ld >= nk, "drb_softmax_rows: bad arguments");
  if (rows == 0) return 0;
  softmax_rows_kernel<<<cdiv(rows, 8), 256, 0, stream>>>(s, rows, nk, ld, scale);
  DRB_LAUNCH_OK();
  return 0;
}
extern "C" int drb_softmax_backward_rows(const float* p, float* dp, long long rows, int nk, int ld, float scale,
                                         cudaStream_t stream) {
  DRB_REQUIRE(p && dp && nk > 0 && ld >= nk, "drb_softmax_backward_rows: bad arguments");
  if (rows == 0) return 0;
  softmax_bwd_rows_kernel<<<cdiv(rows, 8), 256, 0, stream>>>(p, dp, rows, nk, ld, scale);
  DRB_LAUNCH_OK();
  return 0;
}
extern "C" int drb_softmax_weighted_xyz_backward(float* s, int ld, int nq, int nk, const float* xyz, int ld_xyz,
                                                 const float* dcorr, cudaStream_t stream) {
  DRB_REQUIRE(s && xyz && dcorr && nk > 0 && ld >= nk && ld_xyz >= 3, "drb_softmax_weighted_xyz_backward: bad arguments");
  if (nq == 0) return 0;
  corr_bwd_rows_kernel<<<cdiv(nq, 8), 256, 0, stream>>>(s, nq, nk, ld, xyz, ld_xyz, dcorr);
  DRB_LAUNCH_OK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// Batched strided fp32 GEMM on the CUDA cores: C[b](m, n) (+)= alpha * sum_k A[b](m, k) B[b](k, n) with
// arbitrary element strides.  Used for the attention / decoder backward products whose shapes (head
// dim 32, transposed operands, fp32 soft-max matrices) do not fit the 16-bit plane GEMM.
// 64 x 64 tile, 16-deep k step, 256 threads x (4 x 4) outputs.
// ------------------------------------------------------------------------------------------
struct SgemmArgs {
  const float* A; long long a_b, a_m, a_k;
  const float* B; long long b_b, b_k, b_n;
  float* C; long long c_b, c_m;
  int M, N, K;
  float alpha;
  int accumulate;
};
__global__ void __launch_bounds__(256) sgemm_kernel(const SgemmArgs a) {
  __shared__ float As[16][64 + 4];
  __shared__ float Bs[16][64 + 4];
  const int b = blockIdx.z;
  const float* A = a.A + (long long)b * a.a_b;
  const float* B = a.B + (long long)b * a.b_b;
  float* C = a.C + (long long)b * a.c_b;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;     // 16 x 16 threads, each 4 x 4 outputs (strided by 16)
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const bool a_kfast = (a.a_k == 1);
  const bool b_nfast = (a.b_n == 1);
  for (int k0 = 0; k0 < a.K; k0 += 16) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int idx = tid + u * 256;
      int mm, kk;
      if (a_kfast) { kk = idx & 15; mm = idx >> 4; } else { mm = idx & 63; kk = idx >> 6; }
      float v = 0.f;
      if (m0 + mm < a.M && k0 + kk < a.K) v = A[(long long)(m0 + mm) * a.a_m + (long long)(k0 + kk) * a.a_k];
      As[kk][mm] = v;
      int nn, kb;
      if (b_nfast) { nn = idx & 63; kb = idx >> 6; } else { kb = idx & 15; nn = idx >> 4; }
      float w = 0.f;
      if (n0 + nn < a.N && k0 + kb < a.K) w = B[(long long)(k0 + kb) * a.b_k + (long long)(n0 + nn) * a.b_n];
      Bs[kb][nn] = w;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = As[kk][ty + 16 * i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = Bs[kk][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty + 16 * i;
    if (m >= a.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx + 16 * j;
      if (n >= a.N) continue;
      float* dst = C + (long long)m * a.c_m + n;
      const float v = a.alpha * acc[i][j];
      *dst = a.accumulate ? *dst + v : v;
    }
  }
}
// The same product for a long reduction and few output tiles (attention backward: dV = P^T dO, dQ = dS K,
// dK = dS^T Q with N = 32 and K = tokens; decoder backward with N = 256): the 64 x 64 kernel above then runs a
// handful of blocks through a serial, latency-bound k loop (measured 32 us per launch at 500 tokens).  Here a block
// owns a 16 x 32 tile and its 8 warps take interleaved 32-wide k chunks: the A sub-tile goes through a per-warp
// shared-memory tile (broadcast reads), lane = output column, all B values of a chunk are fetched up front, and
// the 8 partial tiles meet in shared memory in warp order (deterministic, no atomics).
static constexpr int kKsM = 16, kKsPitch = 20;
__global__ void __launch_bounds__(256) sgemm_ksplit_kernel(const SgemmArgs a) {
  __shared__ __align__(16) float As[8][32][kKsPitch];     // [warp][k][m]
  __shared__ float red[8][kKsM][32];
  const int b = blockIdx.z;
  const float* __restrict__ A = a.A + (long long)b * a.a_b;
  const float* __restrict__ B = a.B + (long long)b * a.b_b;
  float* C = a.C + (long long)b * a.c_b;
  const int m0 = blockIdx.y * kKsM, n0 = blockIdx.x * 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool a_kfast = (a.a_k == 1);
  const bool n_ok = n0 + lane < a.N;
  float acc[kKsM];
#pragma unroll
  for (int i = 0; i < kKsM; ++i) acc[i] = 0.f;
  for (int k0 = warp * 32; k0 < a.K; k0 += 256) {
    // B values of this chunk (lane = output column): one round of loads in flight
    float bv[32];
#pragma unroll
    for (int kk = 0; kk < 32; ++kk)
      bv[kk] = (n_ok && k0 + kk < a.K) ? B[(long long)(k0 + kk) * a.b_k + (long long)(n0 + lane) * a.b_n] : 0.f;
    __syncwarp();                                        // the previous chunk's tile has been read
    if (a_kfast) {                                       // rows of A are contiguous in k: lane = k
#pragma unroll
      for (int m = 0; m < kKsM; ++m)
        As[warp][lane][m] = (m0 + m < a.M && k0 + lane < a.K) ? A[(long long)(m0 + m) * a.a_m + (k0 + lane)] : 0.f;
    } else {                                             // columns contiguous in m (transposed operand): lane = m
      for (int kk = 0; kk < 32; kk += 2) {
        const int mm = lane & 15, kx = kk + (lane >> 4);
        As[warp][kx][mm] = (m0 + mm < a.M && k0 + kx < a.K)
                               ? A[(long long)(m0 + mm) * a.a_m + (long long)(k0 + kx) * a.a_k] : 0.f;
      }
    }
    __syncwarp();
#pragma unroll
    for (int kk = 0; kk < 32; ++kk) {
      const float4* ar = (const float4*)&As[warp][kk][0];
#pragma unroll
      for (int q = 0; q < kKsM / 4; ++q) {
        const float4 v = ar[q];
        acc[4 * q] = fmaf(v.x, bv[kk], acc[4 * q]);
        acc[4 * q + 1] = fmaf(v.y, bv[kk], acc[4 * q + 1]);
        acc[4 * q + 2] = fmaf(v.z, bv[kk], acc[4 * q + 2]);
        acc[4 * q + 3] = fmaf(v.w, bv[kk], acc[4 * q + 3]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < kKsM; ++i) red[warp][i][lane] = acc[i];
  __syncthreads();
  for (int o = threadIdx.x; o < kKsM * 32; o += 256) {
    const int m = o >> 5, n = o & 31;
    if (m0 + m >= a.M || n0 + n >= a.N) continue;
    float sum = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) sum += red[w][m][n];
    float* dst = C + (long long)(m0 + m) * a.c_m + (n0 + n);
    const float v = a.alpha * sum;
    *dst = a.accumulate ? *dst + v : v;
  }
}

extern "C" int drb_sgemm_strided(const float* A, long long a_b, long long a_m, long long a_k, const float* B,
                                 long long b_b, long long b_k, long long b_n, float* Cm, long long c_b, long long c_m,
                                 int M, int N, int K, int batch, float alpha, int accumulate, cudaStream_t stream) {
  DRB_REQUIRE(A && B && Cm && M >= 0 && N >= 0 && K >= 0 && batch > 0, "drb_sgemm_strided: bad arguments");
  if (M == 0 || N == 0) return 0;
  SgemmArgs a;
  a.A = A; a.a_b = a_b; a.a_m = a_m; a.a_k = a_k;
  a.B = B; a.b_b = b_b; a.b_k = b_k; a.b_n = b_n;
  a.C = Cm; a.c_b = c_b; a.c_m = c_m;
  a.M = M; a.N = N; a.K = K; a.alpha = alpha; a.accumulate = accumulate;
  const long long tiles64 = (long long)cdiv(N, 64) * cdiv(M, 64) * batch;
  if (K >= 128 && tiles64 < 2 * 148) {                   // long reduction, few tiles: in-block split-K
    dim3 grid((unsigned)cdiv(N, 32), (unsigned)cdiv(M, kKsM), (unsigned)batch);
    sgemm_ksplit_kernel<<<grid, 256, 0, stream>>>(a);
    DRB_LAUNCH_OK();
    return 0;
  }
  dim3 grid((unsigned)cdiv(N, 64), (unsigned)cdiv(M, 64), (unsigned)batch);
  sgemm_kernel<<<grid, 256, 0, stream>>>(a);
  DRB_LAUNCH_OK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// Multi-head attention backward (head dim 32) composed from the pieces above.  q / k / v / dout and the
// gradient outputs are [n][ld] row matrices with head h at columns [32 h, 32 h + 32).
// Workspace: 2 * heads * nq * ld_s floats with ld_s = nk rounded up to 4.
//   P  = softmax(scale Q K^T)        dP = dO V^T        dS = P o (dP - rowsum(P o dP)) scale
//   dQ = dS K        dK = dS^T Q        dV = P^T dO
// ------------------------------------------------------------------------------------------
extern "C" size_t drb_mha_backward_workspace_bytes(int nq, int nk, int heads) {
  const long long ld_s = ((long long)nk + 3) / 4 * 4;
  return (size_t)2 * heads * (size_t)nq * (size_t)ld_s * sizeof(float);
}
extern "C" int drb_mha_core_backward(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv,
                                     const float* dout, int ld_do, int nq, int nk, int heads, float scale,
                                     float* dq, int ld_dq, float* dk, int ld_dk, float* dv, int ld_dv,
                                     void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  DRB_REQUIRE(q && k && v && dout && dq && dk && dv && workspace, "drb_mha_core_backward: null argument");
  DRB_REQUIRE(workspace_bytes >= drb_mha_backward_workspace_bytes(nq, nk, heads), "drb_mha_core_backward: workspace too small");
  if (nq == 0 || nk == 0) return 0;
  const long long ld_s = ((long long)nk + 3) / 4 * 4;
  float* P = (float*)workspace;
  float* dP = P + (long long)heads * nq * ld_s;
  const long long sb = (long long)nq * ld_s;
  int rc;
  // S = Q K^T (scale applied inside the soft-max)
  if ((rc = drb_sgemm_strided(q, 32, ldq, 1, k, 32, 1, ldk, P, sb, ld_s, nq, nk, 32, heads, 1.f, 0, stream))) return rc;
  if ((rc = drb_softmax_rows(P, (long long)heads * nq, nk, (int)ld_s, scale, stream))) return rc;
  // dP = dO V^T
  if ((rc = drb_sgemm_strided(dout, 32, ld_do, 1, v, 32, 1, ldv, dP, sb, ld_s, nq, nk, 32, heads, 1.f, 0, stream))) return rc;
  // dV = P^T dO   (before dP is overwritten? independent of dS; P is still the probabilities)
  if ((rc = drb_sgemm_strided(P, sb, 1, ld_s, dout, 32, ld_do, 1, dv, 32, ld_dv, nk, 32, nq, heads, 1.f, 0, stream))) return rc;
  if ((rc = drb_softmax_backward_rows(P, dP, (long long)heads * nq, nk, (int)ld_s, scale, stream))) return rc;
  // dQ = dS K ; dK = dS^T Q
  if ((rc = drb_sgemm_strided(dP, sb, ld_s, 1, k, 32, ldk, 1, dq, 32, ld_dq, nq, 32, nk, heads, 1.f, 0, stream))) return rc;
  if ((rc = drb_sgemm_strided(dP, sb, 1, ld_s, q, 32, ldq, 1, dk, 32, ld_dk, nk, 32, nq, heads, 1.f, 0, stream))) return rc;
  return 0;
}

// ------------------------------------------------------------------------------------------
// Weighted Procrustes backward.  Forward (procrustes.cu): wn = w / max(sum w, 1e-6), weighted centroids
// ca / cb, M = sum wn (a - ca)(b - cb)^T = U S V^T, R = V diag(1, 1, +-1) U^T, t = cb - R ca.
// With P = R M = V (D S) V^T symmetric, dR = Omega R and Omega~_ij = X~_ij / (s_i + s_j) in V's basis,
// X = dM^T R^T - R dM.  The adjoint: A = G_R' R^T, B~_ij = (A~_ij - A~_ji) / (2 (s_i + s_j)),
// dM = R^T B^T - R^T B.  (Checked against torch autograd through torch.linalg.svd, both determinant signs.)
// ------------------------------------------------------------------------------------------
struct ProcBwdArgs {
  const float *a1, *b1, *w1, *a2, *b2, *w2;
  long long a1_ls, b1_ls, w1_ls, a2_ls, b2_ls, w2_ls;
  int n1, n2, ld;
  const float* dpose;       // [layers][3][4]
  float *da1, *db1, *dw1, *da2, *db2, *dw2;   // optional, accumulated; same layer strides as the inputs
};

__device__ double block_sum_bwd(double v, double* sh) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double t = 0.0;
  const int nw = blockDim.x >> 5;
  for (int i = 0; i < nw; ++i) t += sh[i];
  return t;
}

__global__ void __launch_bounds__(256) procrustes_bwd_kernel(ProcBwdArgs p) {
  __shared__ double sh[8];
  __shared__ double s_dM[9], s_dca[3], s_dcb[3], s_ca[3], s_cb[3], s_misc[2];
  const int layer = blockIdx.x;
  const float* a1 = p.a1 + layer * p.a1_ls; const float* b1 = p.b1 + layer * p.b1_ls;
  const float* w1 = p.w1 + layer * p.w1_ls;
  const float* a2 = p.a2 + layer * p.a2_ls; const float* b2 = p.b2 + layer * p.b2_ls;
  const float* w2 = p.w2 + layer * p.w2_ls;
  const int n = p.n1 + p.n2;
  auto fetch = [&](int i, float& w, float a[3], float b[3]) {
    const float *pa, *pb;
    if (i < p.n1) { pa = a1 + (long long)i * p.ld; pb = b1 + (long long)i * p.ld; w = w1[i]; }
    else { const int j = i - p.n1; pa = a2 + (long long)j * p.ld; pb = b2 + (long long)j * p.ld; w = w2[j]; }
    a[0] = pa[0]; a[1] = pa[1]; a[2] = pa[2];
    b[0] = pb[0]; b[1] = pb[1]; b[2] = pb[2];
  };
  double sw = 0, sa[3] = {0, 0, 0}, sb[3] = {0, 0, 0};
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float w, a[3], b[3];
    fetch(i, w, a, b);
    sw += w;
    for (int d = 0; d < 3; ++d) { sa[d] += (double)w * a[d]; sb[d] += (double)w * b[d]; }
  }
  sw = block_sum_bwd(sw, sh);
  for (int d = 0; d < 3; ++d) { sa[d] = block_sum_bwd(sa[d], sh); sb[d] = block_sum_bwd(sb[d], sh); }
  const double denom = sw > 1e-6 ? sw : 1e-6;
  double ca[3], cb[3];
  for (int d = 0; d < 3; ++d) { ca[d] = sa[d] / denom; cb[d] = sb[d] / denom; }
  double cov[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float w, a[3], b[3];
    fetch(i, w, a, b);
    const double wn = (double)w / denom;
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) cov[r * 3 + c] += ((double)a[r] - ca[r]) * (((double)b[c] - cb[c]) * wn);
  }
  __shared__ double scov[9];
  for (int i = 0; i < 9; ++i) {
    const double tot = block_sum_bwd(cov[i], sh);
    if (threadIdx.x == 0) scov[i] = tot;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 0; i < 9; ++i) cov[i] = scov[i];
    double U[9], S[3], V[9];
    svd3(cov, U, S, V);
    double R[9];
    auto build = [&](double sign) {
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c)
          R[r * 3 + c] = V[r * 3 + 0] * U[c * 3 + 0] + V[r * 3 + 1] * U[c * 3 + 1] + sign * V[r * 3 + 2] * U[c * 3 + 2];
    };
    build(1.0);
    const double det = R[0] * (R[4] * R[8] - R[5] * R[7]) - R[1] * (R[3] * R[8] - R[5] * R[6]) +
                       R[2] * (R[3] * R[7] - R[4] * R[6]);
    double sgn = 1.0;
    if (!(det > 0)) { sgn = -1.0; build(-1.0); }
    const float* G = p.dpose + layer * 12;
    double GR[9], gt[3];
    for (int r = 0; r < 3; ++r) {
      gt[r] = G[r * 4 + 3];
      for (int c = 0; c < 3; ++c) GR[r * 3 + c] = G[r * 4 + c];
    }
    // t = cb - R ca
    double dca[3], dcb[3];
    for (int r = 0; r < 3; ++r) {
      dcb[r] = gt[r];
      dca[r] = -(R[0 * 3 + r] * gt[0] + R[1 * 3 + r] * gt[1] + R[2 * 3 + r] * gt[2]);
      for (int c = 0; c < 3; ++c) GR[r * 3 + c] -= gt[r] * ca[c];
    }
    const double s[3] = {S[0], S[1], sgn * S[2]};
    // A = GR R^T ; At = V^T A V
    double A[9], T[9], At[9];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {
        double t = 0;
        for (int k = 0; k < 3; ++k) t += GR[r * 3 + k] * R[c * 3 + k];
        A[r * 3 + c] = t;
      }
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {
        double t = 0;
        for (int k = 0; k < 3; ++k) t += A[r * 3 + k] * V[k * 3 + c];
        T[r * 3 + c] = t;
      }
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {
        double t = 0;
        for (int k = 0; k < 3; ++k) t += V[k * 3 + r] * T[k * 3 + c];
        At[r * 3 + c] = t;
      }
    double Bt[9];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {
        const double den = s[r] + s[c];
        Bt[r * 3 + c] = (r == c || fabs(den) < 1e-300) ? 0.0 : 0.5 * (At[r * 3 + c] - At[c * 3 + r]) / den;
      }
    // B = V Bt V^T
    double B[9];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {
        double t = 0;
        for (int k = 0; k < 3; ++k) t += V[r * 3 + k] * Bt[k * 3 + c];
        T[r * 3 + c] = t;
      }
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {
        double t = 0;
        for (int k = 0; k < 3; ++k) t += T[r * 3 + k] * V[c * 3 + k];
        B[r * 3 + c] = t;
      }
    // dM = R^T B^T - R^T B = R^T (B^T - B)
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {
        double t = 0;
        for (int k = 0; k < 3; ++k) t += R[k * 3 + r] * (B[c * 3 + k] - B[k * 3 + c]);
        s_dM[r * 3 + c] = t;
      }
    // centroid terms: sum_i dac_i = dM cb (1 - sn), sum_i dbc_i = dM^T ca (1 - sn), sn = sum wn
    const double sn = sw / denom;
    for (int r = 0; r < 3; ++r) {
      double t1 = 0, t2 = 0;
      for (int k = 0; k < 3; ++k) { t1 += s_dM[r * 3 + k] * cb[k]; t2 += s_dM[k * 3 + r] * ca[k]; }
      s_dca[r] = dca[r] - t1 * (1.0 - sn);
      s_dcb[r] = dcb[r] - t2 * (1.0 - sn);
      s_ca[r] = ca[r]; s_cb[r] = cb[r];
    }
    s_misc[0] = denom;
    s_misc[1] = sw > 1e-6 ? 1.0 : 0.0;
  }
  __syncthreads();
  double dM[9], dcat[3], dcbt[3];
  for (int i = 0; i < 9; ++i) dM[i] = s_dM[i];
  for (int i = 0; i < 3; ++i) { dcat[i] = s_dca[i]; dcbt[i] = s_dcb[i]; }
  // pass 1: T = sum_k dwn_k w_k
  double tsum = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float w, a[3], b[3];
    fetch(i, w, a, b);
    double dwn = 0.0;
    for (int r = 0; r < 3; ++r) {
      double t = 0;
      for (int c = 0; c < 3; ++c) t += dM[r * 3 + c] * ((double)b[c] - cb[c]);
      dwn += ((double)a[r] - ca[r]) * t + (double)a[r] * dcat[r] + (double)b[r] * dcbt[r];
    }
    tsum += dwn * (double)w;
  }
  tsum = block_sum_bwd(tsum, sh);
  const double corr = s_misc[1] * tsum / (denom * denom);
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float w, a[3], b[3];
    fetch(i, w, a, b);
    const double wn = (double)w / denom;
    double ga[3], gb[3], dwn = 0.0;
    for (int r = 0; r < 3; ++r) {
      double t = 0, u = 0;
      for (int c = 0; c < 3; ++c) {
        t += dM[r * 3 + c] * ((double)b[c] - cb[c]);      // (dM bc)_r
        u += dM[c * 3 + r] * ((double)a[c] - ca[c]);      // (dM^T ac)_r
      }
      ga[r] = wn * (t + dcat[r]);
      gb[r] = wn * (u + dcbt[r]);
      dwn += ((double)a[r] - ca[r]) * t + (double)a[r] * dcat[r] + (double)b[r] * dcbt[r];
    }
    const double gw = dwn / denom - corr;
    float *pda, *pdb, *pdw;
    long long j;
    if (i < p.n1) {
      j = i;
      pda = p.da1 ? p.da1 + layer * p.a1_ls + j * p.ld : nullptr;
      pdb = p.db1 ? p.db1 + layer * p.b1_ls + j * p.ld : nullptr;
      pdw = p.dw1 ? p.dw1 + layer * p.w1_ls + j : nullptr;
    } else {
      j = i - p.n1;
      pda = p.da2 ? p.da2 + layer * p.a2_ls + j * p.ld : nullptr;
      pdb = p.db2 ? p.db2 + layer * p.b2_ls + j * p.ld : nullptr;
      pdw = p.dw2 ? p.dw2 + layer * p.w2_ls + j : nullptr;
    }
    if (pda) { pda[0] += (float)ga[0]; pda[1] += (float)ga[1]; pda[2] += (float)ga[2]; }
    if (pdb) { pdb[0] += (float)gb[0]; pdb[1] += (float)gb[1]; pdb[2] += (float)gb[2]; }
    if (pdw) pdw[0] += (float)gw;
  }
}

extern "C" int drb_procrustes_backward(const float* a1, long long a1_ls, const float* b1, long long b1_ls,
                                       const float* w1, long long w1_ls, int n1, const float* a2, long long a2_ls,
                                       const float* b2, long long b2_ls, const float* w2, long long w2_ls, int n2,
                                       int ld_pts, int layers, const float* dpose, float* da1, float* db1,
                                       float* dw1, float* da2, float* db2, float* dw2, cudaStream_t stream) {
  DRB_REQUIRE(dpose && layers > 0 && ld_pts >= 3, "drb_procrustes_backward: bad arguments");
  DRB_REQUIRE(n1 == 0 || (a1 && b1 && w1), "drb_procrustes_backward: segment 1 pointers");
  DRB_REQUIRE(n2 == 0 || (a2 && b2 && w2), "drb_procrustes_backward: segment 2 pointers");
  DRB_REQUIRE(!(da1 && a1_ls == 0) && !(db1 && b1_ls == 0) && !(da2 && a2_ls == 0) && !(db2 && b2_ls == 0),
              "drb_procrustes_backward: a broadcast operand cannot receive a per-layer gradient");
  ProcBwdArgs p;
  p.a1 = a1; p.b1 = b1; p.w1 = w1; p.a2 = a2; p.b2 = b2; p.w2 = w2;
  p.a1_ls = a1_ls; p.b1_ls = b1_ls; p.w1_ls = w1_ls; p.a2_ls = a2_ls; p.b2_ls = b2_ls; p.w2_ls = w2_ls;
  p.n1 = n1; p.n2 = n2; p.ld = ld_pts; p.dpose = dpose;
  p.da1 = da1; p.db1 = db1; p.dw1 = dw1; p.da2 = da2; p.db2 = db2; p.dw2 = dw2;
  procrustes_bwd_kernel<<<layers, 256, 0, stream>>>(p);
  DRB_LAUNCH_OK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// Voxel-average down-sampling backward (one round): every member row of a cell receives the cell's
// gradient divided by the member count.  sorted_rows / seg_start as recorded by the forward round.
// ------------------------------------------------------------------------------------------
__global__ void segment_mean_bwd_kernel(const float* __restrict__ dout, int ld_out, const int* __restrict__ sorted_rows,
                                        const int* __restrict__ seg_start, int n_seg, int c, float* __restrict__ din,
                                        int ld_in) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= n_seg) return;
  const int s0 = seg_start[warp], s1 = seg_start[warp + 1];
  const float inv = __fdiv_rn(1.f, (float)(s1 - s0));
  for (int cc = lane * 4; cc < c; cc += 128) {
    float4 gv = *(const float4*)(dout + (long long)warp * ld_out + cc);
    gv.x *= inv; gv.y *= inv; gv.z *= inv; gv.w *= inv;
    for (int j = s0; j < s1; ++j) *(float4*)(din + (long long)sorted_rows[j] * ld_in + cc) = gv;
  }
}
extern "C" int drb_segment_mean_backward(const float* dout, int ld_out, const int* sorted_rows, const int* seg_start,
                                         int n_seg, int c, float* din, int ld_in, cudaStream_t stream) {
  DRB_REQUIRE(dout && sorted_rows && seg_start && din && c % 4 == 0 && ld_out % 4 == 0 && ld_in % 4 == 0,
              "drb_segment_mean_backward: bad arguments");
  if (n_seg == 0) return 0;
  segment_mean_bwd_kernel<<<cdiv(n_seg, 8), 256, 0, stream>>>(dout, ld_out, sorted_rows, seg_start, n_seg, c, din, ld_in);
  DRB_LAUNCH_OK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// Fused clip_grad_norm_ + AdamW over a table of tensors (train_nerf_regtr.py:96-102,232-239):
//   norm  = ||all grads||_2 ; coef = min(1, max_norm / (norm + 1e-6))          (torch clip_grad_norm_)
//   p    *= 1 - lr * wd ; m = b1 m + (1 - b1) g ; v = b2 v + (1 - b2) g^2
//   p    -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)              (torch.optim.AdamW)
// No host synchronisation: the norm stays on the device.
// ------------------------------------------------------------------------------------------
struct AdamChunk { int tensor; int pad; long long offset; };
struct AdamTensor { float* p; const float* g; float* m; float* v; long long n; };

}  // namespace drb

struct drb_adamw {
  int n = 0;
  long long total = 0;
  int n_chunks = 0;
  drb::AdamTensor* d_tensors = nullptr;
  drb::AdamChunk* d_chunks = nullptr;
  float* d_state = nullptr;       // exp_avg | exp_avg_sq, flat
  double* d_norm2 = nullptr;
  long long step = 0;
  std::vector<drb::AdamTensor> h_tensors;
};

namespace drb {

static constexpr int kAdamChunk = 16384;

__global__ void grad_norm2_kernel(const AdamTensor* __restrict__ t, const AdamChunk* __restrict__ chunks,
                                  double* __restrict__ norm2) {
  const AdamChunk ch = chunks[blockIdx.x];
  const AdamTensor T = t[ch.tensor];
  long long end = ch.offset + kAdamChunk;
  if (end > T.n) end = T.n;
  double s = 0.0;
  for (long long i = ch.offset + threadIdx.x; i < end; i += blockDim.x) {
    const float g = T.g[i];
    s += (double)g * (double)g;
  }
  s = warp_sum(s);
  __shared__ double sh[8];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t2 = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t2 += sh[i];
    atomicAdd(norm2, t2);
  }
}

__global__ void adamw_kernel(const AdamTensor* __restrict__ t, const AdamChunk* __restrict__ chunks,
                             const double* __restrict__ norm2, float max_norm, float lr, float beta1, float beta2,
                             float eps, float wd, float bc1, float bc2_sqrt) {
  const AdamChunk ch = chunks[blockIdx.x];
  const AdamTensor T = t[ch.tensor];
  long long end = ch.offset + kAdamChunk;
  if (end > T.n) end = T.n;
  float coef = 1.f;
  if (max_norm > 0.f) {
    const float nrm = (float)sqrt(*norm2);
    coef = fminf(1.f, max_norm / (nrm + 1e-6f));
  }
  const float step_size = lr / bc1;
  for (long long i = ch.offset + threadIdx.x; i < end; i += blockDim.x) {
    const float g = T.g[i] * coef;
    float p = T.p[i];
    p *= 1.f - lr * wd;
    const float m = beta1 * T.m[i] + (1.f - beta1) * g;
    const float v = beta2 * T.v[i] + (1.f - beta2) * g * g;
    T.m[i] = m;
    T.v[i] = v;
    const float den = sqrtf(v) / bc2_sqrt + eps;
    T.p[i] = p - step_size * (m / den);
  }
}

}  // namespace drb

using namespace drb;

extern "C" void drb_adamw_destroy(drb_adamw* o);

extern "C" int drb_adamw_create(int n, float* const* params_host, const long long* numels_host, drb_adamw** out) {
  DRB_REQUIRE(n > 0 && params_host && numels_host && out, "drb_adamw_create: bad arguments");
  drb_adamw* o = new drb_adamw();
  o->n = n;
  long long total = 0;
  for (int i = 0; i < n; ++i) {
    if (!params_host[i] || numels_host[i] <= 0) { delete o; set_error("drb_adamw_create: tensor %d is empty", i); return DRB_EINVAL; }
    total += (numels_host[i] + 3) / 4 * 4;
  }
  o->total = total;
  if (cudaMalloc(&o->d_state, sizeof(float) * 2 * (size_t)total) != cudaSuccess ||
      cudaMalloc(&o->d_norm2, sizeof(double)) != cudaSuccess) {
    set_error("drb_adamw_create: out of device memory");
    drb_adamw_destroy(o);
    return DRB_ENOMEM;
  }
  cudaMemset(o->d_state, 0, sizeof(float) * 2 * (size_t)total);
  std::vector<AdamChunk> chunks;
  long long off = 0;
  for (int i = 0; i < n; ++i) {
    AdamTensor t;
    t.p = params_host[i]; t.g = nullptr; t.n = numels_host[i];
    t.m = o->d_state + off; t.v = o->d_state + total + off;
    off += (numels_host[i] + 3) / 4 * 4;
    o->h_tensors.push_back(t);
    for (long long c0 = 0; c0 < t.n; c0 += kAdamChunk) { AdamChunk c; c.tensor = i; c.pad = 0; c.offset = c0; chunks.push_back(c); }
  }
  o->n_chunks = (int)chunks.size();
  if (cudaMalloc(&o->d_tensors, sizeof(AdamTensor) * n) != cudaSuccess ||
      cudaMalloc(&o->d_chunks, sizeof(AdamChunk) * chunks.size()) != cudaSuccess) {
    set_error("drb_adamw_create: out of device memory");
    drb_adamw_destroy(o);
    return DRB_ENOMEM;
  }
  cudaMemcpy(o->d_chunks, chunks.data(), sizeof(AdamChunk) * chunks.size(), cudaMemcpyHostToDevice);
  *out = o;
  return 0;
}

extern "C" void drb_adamw_destroy(drb_adamw* o) {
  if (!o) return;
  cudaFree(o->d_state); cudaFree(o->d_norm2); cudaFree(o->d_tensors); cudaFree(o->d_chunks);
  delete o;
}

extern "C" int drb_adamw_copy_state(drb_adamw* o, int i, float* exp_avg_out, float* exp_avg_sq_out, long long* step,
                                    cudaStream_t stream) {
  DRB_REQUIRE(o && i >= 0 && i < o->n, "drb_adamw_copy_state: bad arguments");
  const AdamTensor& t = o->h_tensors[(size_t)i];
  if (exp_avg_out)
    DRB_CUDA_OK(cudaMemcpyAsync(exp_avg_out, t.m, sizeof(float) * (size_t)t.n, cudaMemcpyDeviceToDevice, stream));
  if (exp_avg_sq_out)
    DRB_CUDA_OK(cudaMemcpyAsync(exp_avg_sq_out, t.v, sizeof(float) * (size_t)t.n, cudaMemcpyDeviceToDevice, stream));
  if (step) *step = o->step;
  return 0;
}

extern "C" int drb_adamw_set_step(drb_adamw* o, long long step) {
  DRB_REQUIRE(o && step >= 0, "drb_adamw_set_step: bad arguments");
  o->step = step;
  return 0;
}

extern "C" int drb_adamw_step(drb_adamw* o, const float* const* grads_host, float lr, float beta1, float beta2,
                              float eps, float weight_decay, float max_grad_norm, cudaStream_t stream) {
  DRB_REQUIRE(o && grads_host, "drb_adamw_step: bad arguments");
  bool changed = false;
  for (int i = 0; i < o->n; ++i) {
    DRB_REQUIRE(grads_host[i] != nullptr, "drb_adamw_step: gradient %d is null", i);
    if (o->h_tensors[i].g != grads_host[i]) { o->h_tensors[i].g = grads_host[i]; changed = true; }
  }
  if (changed)   // pageable -> device: ordered before the kernels below on `stream`
    DRB_CUDA_OK(cudaMemcpyAsync(o->d_tensors, o->h_tensors.data(), sizeof(AdamTensor) * o->n, cudaMemcpyHostToDevice, stream));
  o->step += 1;
  DRB_CUDA_OK(cudaMemsetAsync(o->d_norm2, 0, sizeof(double), stream));
  {   // the norm is always computed: drb_adamw_grad_norm reports it
    grad_norm2_kernel<<<o->n_chunks, 256, 0, stream>>>(o->d_tensors, o->d_chunks, o->d_norm2);
    DRB_LAUNCH_OK();
  }
  const float bc1 = 1.f - powf(beta1, (float)o->step);
  const float bc2 = 1.f - powf(beta2, (float)o->step);
  adamw_kernel<<<o->n_chunks, 256, 0, stream>>>(o->d_tensors, o->d_chunks, o->d_norm2, max_grad_norm, lr, beta1, beta2,
                                                eps, weight_decay, bc1, sqrtf(bc2));
  DRB_LAUNCH_OK();
  return 0;
}

extern "C" int drb_adamw_grad_norm(drb_adamw* o, double* host_norm, cudaStream_t stream) {
  DRB_REQUIRE(o && host_norm, "drb_adamw_grad_norm: bad arguments");
  double n2 = 0.0;
  DRB_CUDA_OK(cudaMemcpyAsync(&n2, o->d_norm2, sizeof(double), cudaMemcpyDeviceToHost, stream));
  DRB_CUDA_OK(cudaStreamSynchronize(stream));
  *host_norm = sqrt(n2);
  return 0;
}
