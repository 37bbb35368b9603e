// HBM-bound helpers around the tensor-core GEMM: plane splitting, weight packing, im2col,
// BatchNorm statistics / apply, max-pool, FPN nearest-upsample + add.  All channels-last.
#include "common.cuh"

#include <math.h>
#include <stdarg.h>

namespace drb {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* drb_last_error(void) { return g_err; }
extern "C" int drb_abi_version(void) { return DRB_ABI_VERSION; }

static inline int grid_for(long long n, int block, int cap = 148 * 16) {
  long long g = (n + block - 1) / block;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// ------------------------------------------------------------------------------------------
__global__ void split_kernel(const float* __restrict__ x, plane_t* __restrict__ hi,
                             plane_t* __restrict__ lo, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    plane_t h, l;
    split16(x[i], lo != nullptr, h, l);
    hi[i] = h;
    if (lo) lo[i] = l;
  }
}

extern "C" int drb_split_planes(const float* x, void* hi, void* lo, long long n,
                                cudaStream_t stream) {
  DRB_REQUIRE(x && hi && n >= 0, "drb_split_planes: bad arguments");
  if (n == 0) return 0;
  split_kernel<<<grid_for(n, 256), 256, 0, stream>>>(x, (plane_t*)hi, (plane_t*)lo, n);
  DRB_LAUNCH_OK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// weights [cout][cin][taps] -> [taps][cout][cin_pad]
__global__ void pack_w_kernel(const float* __restrict__ w, int cout, int cin, int taps, int cin_pad,
                              float scale, plane_t* __restrict__ hi, plane_t* __restrict__ lo) {
  const long long total = (long long)taps * cout * cin_pad;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c = (int)(i % cin_pad);
    const int o = (int)((i / cin_pad) % cout);
    const int t = (int)(i / ((long long)cin_pad * cout));
    const float v = (c < cin) ? w[((long long)o * cin + c) * taps + t] * scale : 0.f;
    plane_t h, l;
    split16(v, lo != nullptr, h, l);
    hi[i] = h;
    if (lo) lo[i] = l;
  }
}
// weights [cout][cin][taps] -> [cout][kpad], k = tap*cin + c
__global__ void pack_w_im2col_kernel(const float* __restrict__ w, int cout, int cin, int taps,
                                     int kpad, float scale, plane_t* __restrict__ hi,
                                     plane_t* __restrict__ lo) {
  const long long total = (long long)cout * kpad;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int k = (int)(i % kpad);
    const int o = (int)(i / kpad);
    float v = 0.f;
    if (k < taps * cin) {
      const int t = k / cin, c = k % cin;
      v = w[((long long)o * cin + c) * taps + t] * scale;
    }
    plane_t h, l;
    split16(v, lo != nullptr, h, l);
    hi[i] = h;
    if (lo) lo[i] = l;
  }
}

extern "C" int drb_pack_conv_weight(const float* w, int cout, int cin, int taps, int cin_pad,
                                    float scale, void* hi, void* lo, cudaStream_t stream) {
  DRB_REQUIRE(w && hi && cout > 0 && cin > 0 && taps > 0 && cin_pad >= cin,
              "drb_pack_conv_weight: bad arguments");
  pack_w_kernel<<<grid_for((long long)taps * cout * cin_pad, 256), 256, 0, stream>>>(
      w, cout, cin, taps, cin_pad, scale == 0.f ? 1.f : scale, (plane_t*)hi, (plane_t*)lo);
  DRB_LAUNCH_OK();
  return 0;
}
extern "C" int drb_pack_conv_weight_im2col(const float* w, int cout, int cin, int taps, int kpad,
                                           float scale, void* hi, void* lo, cudaStream_t stream) {
  DRB_REQUIRE(w && hi && cout > 0 && cin > 0 && taps > 0 && kpad >= cin * taps,
              "drb_pack_conv_weight_im2col: bad arguments");
  pack_w_im2col_kernel<<<grid_for((long long)cout * kpad, 256), 256, 0, stream>>>(
      w, cout, cin, taps, kpad, scale == 0.f ? 1.f : scale, (plane_t*)hi, (plane_t*)lo);
  DRB_LAUNCH_OK();
  return 0;
}

// max |w| of a tensor (weight pre-scale selection); result accumulates into *out (caller zeroes it).
__global__ void absmax_kernel(const float* __restrict__ w, long long n, unsigned int* __restrict__ out) {
  float m = 0.f;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    m = fmaxf(m, fabsf(w[i]));
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));   // non-negative floats order as uints
}

// Power-of-two scale that maps max|w| into [256, 512): keeps hi and lo = fp16(x - hi) in fp16's
// normal range with head-room against overflow.  Synchronous (one 4-byte read back).
extern "C" int drb_weight_scale(const float* w, long long n, float* host_scale, cudaStream_t stream) {
  DRB_REQUIRE(w && host_scale && n > 0, "drb_weight_scale: bad arguments");
  unsigned int* d = nullptr;
  DRB_CUDA_OK(cudaMallocAsync(&d, sizeof(unsigned int), stream));
  DRB_CUDA_OK(cudaMemsetAsync(d, 0, sizeof(unsigned int), stream));
  absmax_kernel<<<grid_for(n, 256, 296), 256, 0, stream>>>(w, n, d);
  unsigned int bits = 0;
  DRB_CUDA_OK(cudaMemcpyAsync(&bits, d, sizeof(bits), cudaMemcpyDeviceToHost, stream));
  DRB_CUDA_OK(cudaStreamSynchronize(stream));
  DRB_CUDA_OK(cudaFreeAsync(d, stream));
  float mx;
  memcpy(&mx, &bits, sizeof(mx));
  float scale = 1.f;
  if (mx > 0.f && isfinite(mx)) {
    int e = 0;
    frexpf(mx, &e);                 // mx = f * 2^e, f in [0.5, 1)
    scale = ldexpf(1.f, 9 - e);     // mx * scale in [256, 512)
  }
  *host_scale = scale;
  return 0;
}

// ------------------------------------------------------------------------------------------
// im2col: one thread per (output voxel, k) element; k fastest so that writes coalesce.
__global__ void im2col_kernel(drb_im2col_desc d, int od, int oh, int ow, plane_t* __restrict__ hi,
                              plane_t* __restrict__ lo) {
  const long long rows = (long long)d.g * od * oh * ow;
  const long long total = rows * d.kpad;
  const int kk = d.k * d.k * d.k * d.c;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int k = (int)(i % d.kpad);
    long long r = i / d.kpad;
    float v = 0.f;
    if (k < kk) {
      const int ch = k % d.c;
      int t = k / d.c;
      const int kx = t % d.k; t /= d.k;
      const int ky = t % d.k; t /= d.k;
      const int kz = t;
      const int x = (int)(r % ow); r /= ow;
      const int y = (int)(r % oh); r /= oh;
      const int z = (int)(r % od); r /= od;
      const int g = (int)r;
      const int iz = z * d.stride - d.pad + kz;
      const int iy = y * d.stride - d.pad + ky;
      const int ix = x * d.stride - d.pad + kx;
      if (iz >= 0 && iz < d.d && iy >= 0 && iy < d.h && ix >= 0 && ix < d.w)
        v = d.x[g * d.sg + ch * d.sc + iz * d.sd + iy * d.sh + ix * d.sw];
    }
    plane_t h, l;
    split16(v, lo != nullptr, h, l);
    hi[i] = h;
    if (lo) lo[i] = l;
  }
}

// Channels-last input with a multiple of 8 channels and no padding columns (every strided convolution of the
// backbone): one thread per (output voxel, tap, 8 channels) - 32 contiguous bytes in, 16 contiguous bytes out per
// plane, 32-bit index arithmetic.  (The element-per-thread kernel above spends ~100 instructions and two 64-bit
// divisions on every 2-byte store: 275 us for layer2.0.conv2 at 128^3, 0.4 TB/s.)
__global__ void __launch_bounds__(256) im2col_cl8_kernel(drb_im2col_desc d, int od, int oh, int ow,
                                                         plane_t* __restrict__ hi, plane_t* __restrict__ lo) {
  const int cg = d.c >> 3;                       // 8-channel groups per tap
  const int taps = d.k * d.k * d.k;
  const int per_row = taps * cg;
  const long long rows = (long long)d.g * od * oh * ow;
  const long long total = rows * per_row;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const bool pair = lo != nullptr;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long row = i / per_row;
    const int rem = (int)(i - row * per_row);
    const int tap = rem / cg, c8 = (rem - tap * cg) << 3;
    const int kx = tap % d.k, ky = (tap / d.k) % d.k, kz = tap / (d.k * d.k);
    int r = (int)row;                            // rows < 2^31 (checked by the host)
    const int x = r % ow; r /= ow;
    const int y = r % oh; r /= oh;
    const int z = r % od; r /= od;
    const int g = r;
    const int iz = z * d.stride - d.pad + kz, iy = y * d.stride - d.pad + ky, ix = x * d.stride - d.pad + kx;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (iz >= 0 && iz < d.d && iy >= 0 && iy < d.h && ix >= 0 && ix < d.w) {
      const float* src = d.x + g * d.sg + iz * d.sd + iy * d.sh + ix * d.sw + c8;
      a = *(const float4*)src;
      b = *(const float4*)(src + 4);
    }
    const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint32_t hw[4], lw[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      plane_t h0, l0, h1, l1;
      split16(v[2 * u], pair, h0, l0);
      split16(v[2 * u + 1], pair, h1, l1);
      hw[u] = pack16x2(h0, h1);
      lw[u] = pack16x2(l0, l1);
    }
    const long long o = row * d.kpad + (long long)tap * d.c + c8;
    *(uint4*)(hi + o) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
    if (pair) *(uint4*)(lo + o) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
  }
}

extern "C" int drb_im2col(const drb_im2col_desc* d, void* hi, void* lo, cudaStream_t stream) {
  DRB_REQUIRE(d && d->x && hi, "drb_im2col: null argument");
  DRB_REQUIRE(d->k >= 1 && d->stride >= 1 && d->kpad >= d->k * d->k * d->k * d->c &&
                  d->kpad % 64 == 0,
              "drb_im2col: bad kernel/kpad");
  const int od = (d->d + 2 * d->pad - d->k) / d->stride + 1;
  const int oh = (d->h + 2 * d->pad - d->k) / d->stride + 1;
  const int ow = (d->w + 2 * d->pad - d->k) / d->stride + 1;
  const long long total = (long long)d->g * od * oh * ow * d->kpad;
  const long long rows = (long long)d->g * od * oh * ow;
  const bool aligned = ((uintptr_t)d->x & 15) == 0 && ((uintptr_t)hi & 15) == 0 && (!lo || ((uintptr_t)lo & 15) == 0) &&
                       d->sw % 4 == 0 && d->sh % 4 == 0 && d->sd % 4 == 0 && d->sg % 4 == 0;
  if (d->sc == 1 && d->c % 8 == 0 && d->kpad == d->k * d->k * d->k * d->c && rows < (1LL << 31) && aligned) {
    const long long groups = total / 8;
    im2col_cl8_kernel<<<grid_for(groups, 256, 148 * 32), 256, 0, stream>>>(*d, od, oh, ow, (plane_t*)hi, (plane_t*)lo);
    DRB_LAUNCH_OK();
    return 0;
  }
  im2col_kernel<<<grid_for(total, 256, 148 * 32), 256, 0, stream>>>(*d, od, oh, ow, (plane_t*)hi,
                                                                    (plane_t*)lo);
  DRB_LAUNCH_OK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// Stem fast path (conv1 5^3 s2 over the 4 rgba channels of the caller's strided [1,7,Z,X,Y] view):
// (1) pack channels 3..6 into a compact channels-last fp32 volume [d][h][w][4], (2) im2col from it
// with one warp per output voxel: lane j owns taps 4j..4j+3 (16 consecutive k), float4 reads,
// 32-byte plane writes (1 KB contiguous per warp and plane).
__global__ void pack_rgba_kernel(const float* __restrict__ x, long long sc, long long sd, long long sh,
                                 long long sw, int d, int h, int w, float4* __restrict__ out) {
  const long long total = (long long)d * h * w;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int xw = (int)(i % w);
    const int yh = (int)((i / w) % h);
    const int zd = (int)(i / ((long long)w * h));
    const float* p = x + zd * sd + yh * sh + xw * sw;
    out[i] = make_float4(p[0], p[sc], p[2 * sc], p[3 * sc]);
  }
}

__global__ void im2col_stem_kernel(const float4* __restrict__ x, int g, int d, int h, int w, int od,
                                   int oh, int ow, plane_t* __restrict__ hi, plane_t* __restrict__ lo) {
  const long long rows = (long long)g * od * oh * ow;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  long long r = warp;
  const int ox = (int)(r % ow); r /= ow;
  const int oy = (int)(r % oh); r /= oh;
  const int oz = (int)(r % od); r /= od;
  const int gi = (int)r;
  const bool pair = lo != nullptr;
  uint32_t hw[8], lw[8];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int tap = lane * 4 + t;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tap < 125) {
      const int kx = tap % 5, ky = (tap / 5) % 5, kz = tap / 25;
      const int iz = oz * 2 - 2 + kz, iy = oy * 2 - 2 + ky, ix = ox * 2 - 2 + kx;
      if (iz >= 0 && iz < d && iy >= 0 && iy < h && ix >= 0 && ix < w)
        v = x[(((long long)gi * d + iz) * h + iy) * w + ix];
    }
    plane_t h0, l0, h1, l1, h2, l2, h3, l3;
    split16(v.x, pair, h0, l0); split16(v.y, pair, h1, l1);
    split16(v.z, pair, h2, l2); split16(v.w, pair, h3, l3);
    hw[2 * t] = pack16x2(h0, h1); hw[2 * t + 1] = pack16x2(h2, h3);
    lw[2 * t] = pack16x2(l0, l1); lw[2 * t + 1] = pack16x2(l2, l3);
  }
  const long long off = warp * 512 + lane * 16;
  *(uint4*)(hi + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
  *(uint4*)(hi + off + 8) = make_uint4(hw[4], hw[5], hw[6], hw[7]);
  if (pair) {
    *(uint4*)(lo + off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
    *(uint4*)(lo + off + 8) = make_uint4(lw[4], lw[5], lw[6], lw[7]);
  }
}

// x: fp32 view of ONE grid with element strides (sc, sd, sh, sw) already offset to the first of the 4
// channels; scratch: d*h*w float4; planes [do][ho][wo][512].
extern "C" int drb_im2col_stem(const float* x, long long sc, long long sd, long long sh, long long sw,
                               int d, int h, int w, void* scratch, void* hi, void* lo,
                               cudaStream_t stream) {
  DRB_REQUIRE(x && scratch && hi, "drb_im2col_stem: null argument");
  const long long vox = (long long)d * h * w;
  pack_rgba_kernel<<<grid_for(vox, 256, 148 * 32), 256, 0, stream>>>(x, sc, sd, sh, sw, d, h, w,
                                                                    (float4*)scratch);
  DRB_LAUNCH_OK();
  const int od = (d + 4 - 5) / 2 + 1, oh = (h + 4 - 5) / 2 + 1, ow = (w + 4 - 5) / 2 + 1;
  const long long rows = (long long)od * oh * ow;
  im2col_stem_kernel<<<cdiv(rows * 32, 256), 256, 0, stream>>>((const float4*)scratch, 1, d, h, w, od, oh,
                                                              ow, (plane_t*)hi, (plane_t*)lo);
  DRB_LAUNCH_OK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// BatchNorm statistics: x [g][m][c].  Block = 64 channels (16 quads of 4, one 16-byte load per thread and row) x
// 16 row lanes, four rows in flight per thread; double accumulation; one fp64 atomic per (block, channel, moment).
// (Round 1 read one float per thread and row: 1.7 - 3.2 TB/s on tensors the producing GEMM had just left in L2.)
__global__ void __launch_bounds__(256) bn_stats_kernel(const float* __restrict__ x, long long m, int c, int rows_per_block,
                                                       double* __restrict__ accum) {
  const int g = blockIdx.z;
  const int c0 = blockIdx.y * 64;
  const int q = threadIdx.x & 15;          // channel quad
  const int rl = threadIdx.x >> 4;         // row lane 0..15
  const int ch = c0 + 4 * q;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > m) r1 = m;
  double s[4] = {0.0, 0.0, 0.0, 0.0}, ss[4] = {0.0, 0.0, 0.0, 0.0};
  if (ch < c) {                            // c % 4 == 0: a quad is in or out as a whole
    const float* base = x + ((long long)g * m) * c + ch;
    long long r = r0 + rl;
    for (; r + 48 < r1; r += 64) {
      const float4 v0 = *(const float4*)(base + r * c);
      const float4 v1 = *(const float4*)(base + (r + 16) * c);
      const float4 v2 = *(const float4*)(base + (r + 32) * c);
      const float4 v3 = *(const float4*)(base + (r + 48) * c);
      const float4 vv[4] = {v0, v1, v2, v3};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        s[0] += (double)vv[k].x; ss[0] += (double)vv[k].x * (double)vv[k].x;
        s[1] += (double)vv[k].y; ss[1] += (double)vv[k].y * (double)vv[k].y;
        s[2] += (double)vv[k].z; ss[2] += (double)vv[k].z * (double)vv[k].z;
        s[3] += (double)vv[k].w; ss[3] += (double)vv[k].w * (double)vv[k].w;
      }
    }
    for (; r < r1; r += 16) {
      const float4 v = *(const float4*)(base + r * c);
      s[0] += (double)v.x; ss[0] += (double)v.x * (double)v.x;
      s[1] += (double)v.y; ss[1] += (double)v.y * (double)v.y;
      s[2] += (double)v.z; ss[2] += (double)v.z * (double)v.z;
      s[3] += (double)v.w; ss[3] += (double)v.w * (double)v.w;
    }
  }
  __shared__ double sh[2][16][65];
#pragma unroll
  for (int j = 0; j < 4; ++j) { sh[0][rl][4 * q + j] = s[j]; sh[1][rl][4 * q + j] = ss[j]; }
  __syncthreads();
  if (threadIdx.x < 128) {
    const int which = threadIdx.x >> 6, cl = threadIdx.x & 63;
    if (c0 + cl < c) {
      double t = 0.0;
#pragma unroll
      for (int k = 0; k < 16; ++k) t += sh[which][k][cl];
      atomicAdd(&accum[((long long)g * c + c0 + cl) * 2 + which], t);
    }
  }
}

extern "C" int drb_bn_stats(const float* x, int g, long long m, int c, double* accum,
                            cudaStream_t stream) {
  DRB_REQUIRE(x && accum && g > 0 && m > 0 && c > 0 && c % 4 == 0, "drb_bn_stats: bad arguments (c must be a multiple of 4)");
  DRB_CUDA_OK(cudaMemsetAsync(accum, 0, sizeof(double) * 2 * g * c, stream));
  int rows_per_block = 256;
  while ((m + rows_per_block - 1) / rows_per_block > 4096) rows_per_block *= 2;
  dim3 grid((unsigned)((m + rows_per_block - 1) / rows_per_block), (unsigned)((c + 63) / 64),
            (unsigned)g);
  bn_stats_kernel<<<grid, 256, 0, stream>>>(x, m, c, rows_per_block, accum);
  DRB_LAUNCH_OK();
  return 0;
}

__global__ void bn_finalize_kernel(const double* __restrict__ accum, int g, long long m, int c,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* running_mean, float* running_var, int training,
                                   float momentum, float eps, float* __restrict__ scale,
                                   float* __restrict__ shift) {
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= c) return;
  const float ga = gamma ? gamma[ch] : 1.f;
  const float be = beta ? beta[ch] : 0.f;
  if (training) {
    float rm = running_mean ? running_mean[ch] : 0.f;
    float rv = running_var ? running_var[ch] : 1.f;
    for (int gi = 0; gi < g; ++gi) {
      const double s = accum[((long long)gi * c + ch) * 2 + 0];
      const double ss = accum[((long long)gi * c + ch) * 2 + 1];
      const double mean = s / (double)m;
      double var = ss / (double)m - mean * mean;
      if (var < 0.0) var = 0.0;
      const float rstd = (float)(1.0 / sqrt(var + (double)eps));
      const float sc = ga * rstd;
      scale[(long long)gi * c + ch] = sc;
      shift[(long long)gi * c + ch] = be - (float)mean * sc;
      const double unbiased = m > 1 ? var * ((double)m / (double)(m - 1)) : var;
      rm = (1.f - momentum) * rm + momentum * (float)mean;
      rv = (1.f - momentum) * rv + momentum * (float)unbiased;
    }
    if (running_mean) running_mean[ch] = rm;
    if (running_var) running_var[ch] = rv;
  } else {
    const float rstd = 1.f / sqrtf(running_var[ch] + eps);
    const float sc = ga * rstd;
    for (int gi = 0; gi < g; ++gi) {
      scale[(long long)gi * c + ch] = sc;
      shift[(long long)gi * c + ch] = be - running_mean[ch] * sc;
    }
  }
}

extern "C" int drb_bn_finalize(const double* accum, int g, long long m, int c, const float* gamma,
                               const float* beta, float* running_mean, float* running_var,
                               int training, float momentum, float eps, float* scale, float* shift,
                               cudaStream_t stream) {
  DRB_REQUIRE(scale && shift && g > 0 && c > 0, "drb_bn_finalize: bad arguments");
  DRB_REQUIRE(training ? accum != nullptr : (running_mean && running_var),
              "drb_bn_finalize: missing statistics source");
  bn_finalize_kernel<<<(c + 127) / 128, 128, 0, stream>>>(accum, g, m, c, gamma, beta, running_mean,
                                                          running_var, training, momentum, eps,
                                                          scale, shift);
  DRB_LAUNCH_OK();
  return 0;
}

// y = relu?(x*scale + shift + residual); 4 channels per thread (c % 4 == 0).
__global__ void scale_shift_act_kernel(const float* __restrict__ x, const float* __restrict__ scale,
                                       const float* __restrict__ shift,
                                       const float* __restrict__ residual, int relu, long long m,
                                       int c, long long total4, float* __restrict__ out,
                                       plane_t* __restrict__ out_hi, plane_t* __restrict__ out_lo) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const int c4 = c >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += stride) {
    const int cc = (int)(i % c4) * 4;
    const long long row = i / c4;
    const int g = (int)(row / m);
    float4 v = *(const float4*)(x + i * 4);
    if (scale) {
      const float4 sc = *(const float4*)(scale + (long long)g * c + cc);
      const float4 sf = *(const float4*)(shift + (long long)g * c + cc);
      v.x = fmaf(v.x, sc.x, sf.x); v.y = fmaf(v.y, sc.y, sf.y);
      v.z = fmaf(v.z, sc.z, sf.z); v.w = fmaf(v.w, sc.w, sf.w);
    }
    if (residual) {
      const float4 r = *(const float4*)(residual + i * 4);
      v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
    }
    if (relu) {
      v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
    }
    if (out) *(float4*)(out + i * 4) = v;
    if (out_hi) {
      const bool pair = out_lo != nullptr;
      plane_t h0, l0, h1, l1, h2, l2, h3, l3;
      split16(v.x, pair, h0, l0); split16(v.y, pair, h1, l1);
      split16(v.z, pair, h2, l2); split16(v.w, pair, h3, l3);
      *(uint2*)(out_hi + i * 4) = make_uint2(pack16x2(h0, h1), pack16x2(h2, h3));
      if (pair) *(uint2*)(out_lo + i * 4) = make_uint2(pack16x2(l0, l1), pack16x2(l2, l3));
    }
  }
}

extern "C" int drb_scale_shift_act(const float* x, const float* scale, const float* shift,
                                   const float* residual, int relu, int g, long long m, int c,
                                   float* out, void* out_hi, void* out_lo, cudaStream_t stream) {
  DRB_REQUIRE(x && g > 0 && m > 0 && c > 0 && c % 4 == 0, "drb_scale_shift_act: bad arguments");
  DRB_REQUIRE((scale == nullptr) == (shift == nullptr), "drb_scale_shift_act: scale/shift pair");
  const long long total4 = (long long)g * m * c / 4;
  scale_shift_act_kernel<<<grid_for(total4, 256, 148 * 32), 256, 0, stream>>>(
      x, scale, shift, residual, relu, m, c, total4, out, (plane_t*)out_hi, (plane_t*)out_lo);
  DRB_LAUNCH_OK();
  return 0;
}

// Fused finalize + apply: every block derives scale / shift of all (g, c) from the fp64 sums (or the
// running statistics) into shared memory, block 0 also performs the running-statistics update; then
// y = relu?(x * scale + shift + residual).  One launch instead of bn_finalize + scale_shift_act.
__global__ void bn_apply_kernel(const float* __restrict__ x, const double* __restrict__ accum, int g, long long m,
                                int c, const float* __restrict__ gamma, const float* __restrict__ beta,
                                float* running_mean, float* running_var, int training, float momentum, float eps,
                                const float* __restrict__ residual, int relu, long long total4,
                                float* __restrict__ out, plane_t* __restrict__ out_hi,
                                plane_t* __restrict__ out_lo) {
  extern __shared__ float s_ss[];          // [g*c] scale, [g*c] shift
  float* s_scale = s_ss;
  float* s_shift = s_ss + (long long)g * c;
  for (int i = threadIdx.x; i < g * c; i += blockDim.x) {
    const int ch = i % c;
    const float ga = gamma ? gamma[ch] : 1.f, be = beta ? beta[ch] : 0.f;
    float sc, sf;
    if (training) {
      const double mean = accum[(long long)i * 2] / (double)m;
      double var = accum[(long long)i * 2 + 1] / (double)m - mean * mean;
      if (var < 0.0) var = 0.0;
      sc = ga * (float)(1.0 / sqrt(var + (double)eps));
      sf = be - (float)mean * sc;
    } else {
      sc = ga * (1.f / sqrtf(running_var[ch] + eps));
      sf = be - running_mean[ch] * sc;
    }
    s_scale[i] = sc;
    s_shift[i] = sf;
  }
  if (training && blockIdx.x == 0 && running_mean && running_var) {
    for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {     // sequential over g: g successive forward calls
      float rm = running_mean[ch], rv = running_var[ch];
      for (int gi = 0; gi < g; ++gi) {
        const double mean = accum[((long long)gi * c + ch) * 2] / (double)m;
        double var = accum[((long long)gi * c + ch) * 2 + 1] / (double)m - mean * mean;
        if (var < 0.0) var = 0.0;
        const double unbiased = m > 1 ? var * ((double)m / (double)(m - 1)) : var;
        rm = (1.f - momentum) * rm + momentum * (float)mean;
        rv = (1.f - momentum) * rv + momentum * (float)unbiased;
      }
      running_mean[ch] = rm;
      running_var[ch] = rv;
    }
  }
  __syncthreads();
  const long long stride = (long long)gridDim.x * blockDim.x;
  const int c4 = c >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += stride) {
    const int cc = (int)(i % c4) * 4;
    const int gi = (int)((i / c4) / m);
    const float* sc = s_scale + (long long)gi * c + cc;
    const float* sf = s_shift + (long long)gi * c + cc;
    float4 v = *(const float4*)(x + i * 4);
    v.x = fmaf(v.x, sc[0], sf[0]); v.y = fmaf(v.y, sc[1], sf[1]);
    v.z = fmaf(v.z, sc[2], sf[2]); v.w = fmaf(v.w, sc[3], sf[3]);
    if (residual) {
      const float4 r = *(const float4*)(residual + i * 4);
      v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
    }
    if (relu) {
      v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
    }
    if (out) *(float4*)(out + i * 4) = v;
    if (out_hi) {
      const bool pair = out_lo != nullptr;
      plane_t h0, l0, h1, l1, h2, l2, h3, l3;
      split16(v.x, pair, h0, l0); split16(v.y, pair, h1, l1);
      split16(v.z, pair, h2, l2); split16(v.w, pair, h3, l3);
      *(uint2*)(out_hi + i * 4) = make_uint2(pack16x2(h0, h1), pack16x2(h2, h3));
      if (pair) *(uint2*)(out_lo + i * 4) = make_uint2(pack16x2(l0, l1), pack16x2(l2, l3));
    }
  }
}

extern "C" int drb_bn_apply(const float* x, const double* accum, int g, long long m, int c, const float* gamma,
                            const float* beta, float* running_mean, float* running_var, int training,
                            float momentum, float eps, const float* residual, int relu, float* out, void* out_hi,
                            void* out_lo, cudaStream_t stream) {
  DRB_REQUIRE(x && g > 0 && m > 0 && c > 0 && c % 4 == 0 && (out || out_hi), "drb_bn_apply: bad arguments");
  DRB_REQUIRE(training ? accum != nullptr : (running_mean && running_var), "drb_bn_apply: missing statistics");
  const size_t smem = sizeof(float) * 2 * (size_t)g * c;
  DRB_REQUIRE(smem <= 48 * 1024, "drb_bn_apply: g*c = %d exceeds the shared-memory table", g * c);
  const long long total4 = (long long)g * m * c / 4;
  // few, fat blocks: every block recomputes the (g, c) table
  int grid = (int)((total4 + 256 * 8 - 1) / (256 * 8));
  if (grid > 148 * 8) grid = 148 * 8;
  if (grid < 1) grid = 1;
  bn_apply_kernel<<<grid, 256, smem, stream>>>(x, accum, g, m, c, gamma, beta, running_mean, running_var, training,
                                               momentum, eps, residual, relu, total4, out, (plane_t*)out_hi,
                                               (plane_t*)out_lo);
  DRB_LAUNCH_OK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// Small tensors (the deep ResNet stages: m = 512 ... 8 voxels per grid, 256 ... 2048 channels): statistics,
// running-buffer update, normalisation (+ residual, ReLU, plane split) and the backward's mean / rstd / scale /
// shift in ONE launch.  The three launches of the general path (memset + bn_stats + bn_apply, + bn_save in grad
// mode) cost ~30 us per BatchNorm there, all of it latency (46 of the 53 BatchNorms of a forward).
// A block owns 32 channels of all g grids: 8 channel quads x 32 row lanes, fp64 sums as in bn_stats_kernel,
// grids in order (the running buffers see g successive forward calls, like bn_apply_kernel).
// ------------------------------------------------------------------------------------------
static constexpr int kBnSmallMaxRows = 1024;
__global__ void __launch_bounds__(256) bn_small_kernel(const float* __restrict__ x, int g, int m, int c,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       float* running_mean, float* running_var, int training,
                                                       float momentum, float eps, const float* __restrict__ residual,
                                                       int relu, float* __restrict__ out, plane_t* __restrict__ out_hi,
                                                       plane_t* __restrict__ out_lo, float* __restrict__ save_mean,
                                                       float* __restrict__ save_rstd, float* __restrict__ save_scale,
                                                       float* __restrict__ save_shift) {
  __shared__ double s_sum[32][33], s_sq[32][33];       // [row lane][channel]
  __shared__ float s_scale[32], s_shift[32];
  const int c0 = blockIdx.x * 32;
  const int q = threadIdx.x & 7, rl = threadIdx.x >> 3;            // channel quad, row lane
  const int ch4 = c0 + 4 * q;
  const bool ch_ok = ch4 < c;                                      // c % 4 == 0: a quad is in or out as a whole
  for (int gi = 0; gi < g; ++gi) {
    const float* xg = x + (long long)gi * m * c;
    if (training) {
      double s[4] = {0.0, 0.0, 0.0, 0.0}, ss[4] = {0.0, 0.0, 0.0, 0.0};
      if (ch_ok) {
        // four rows in flight per thread: a block is alone on its SM with 8 warps, one load at a time per
        // thread made every pass a chain of L2 round trips (25 us per launch, ncu)
        int r = rl;
        for (; r + 96 < m; r += 128) {
          float4 vv[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) vv[k] = *(const float4*)(xg + (long long)(r + 32 * k) * c + ch4);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            s[0] += (double)vv[k].x; ss[0] += (double)vv[k].x * (double)vv[k].x;
            s[1] += (double)vv[k].y; ss[1] += (double)vv[k].y * (double)vv[k].y;
            s[2] += (double)vv[k].z; ss[2] += (double)vv[k].z * (double)vv[k].z;
            s[3] += (double)vv[k].w; ss[3] += (double)vv[k].w * (double)vv[k].w;
          }
        }
        for (; r < m; r += 32) {
          const float4 v = *(const float4*)(xg + (long long)r * c + ch4);
          s[0] += (double)v.x; ss[0] += (double)v.x * (double)v.x;
          s[1] += (double)v.y; ss[1] += (double)v.y * (double)v.y;
          s[2] += (double)v.z; ss[2] += (double)v.z * (double)v.z;
          s[3] += (double)v.w; ss[3] += (double)v.w * (double)v.w;
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) { s_sum[rl][4 * q + j] = s[j]; s_sq[rl][4 * q + j] = ss[j]; }
    }
    __syncthreads();
    if (threadIdx.x < 32) {
      const int ch = c0 + threadIdx.x;
      if (ch < c) {
        const float ga = gamma ? gamma[ch] : 1.f, be = beta ? beta[ch] : 0.f;
        float mu, rs;
        if (training) {
          double a = 0.0, b = 0.0;
          for (int r = 0; r < 32; ++r) { a += s_sum[r][threadIdx.x]; b += s_sq[r][threadIdx.x]; }
          const double mean = a / (double)m;
          double var = b / (double)m - mean * mean;
          if (var < 0.0) var = 0.0;
          rs = (float)(1.0 / sqrt(var + (double)eps));
          mu = (float)mean;
          if (running_mean && running_var) {
            const double unbiased = m > 1 ? var * ((double)m / (double)(m - 1)) : var;
            running_mean[ch] = (1.f - momentum) * running_mean[ch] + momentum * (float)mean;
            running_var[ch] = (1.f - momentum) * running_var[ch] + momentum * (float)unbiased;
          }
        } else {
          rs = 1.f / sqrtf(running_var[ch] + eps);
          mu = running_mean[ch];
        }
        const float sc = ga * rs, sf = be - mu * sc;
        s_scale[threadIdx.x] = sc; s_shift[threadIdx.x] = sf;
        if (save_mean) {
          const long long i = (long long)gi * c + ch;
          save_mean[i] = mu; save_rstd[i] = rs; save_scale[i] = sc; save_shift[i] = sf;
        }
      }
    }
    __syncthreads();
    if (ch_ok) {
      const float4 sc = *(const float4*)&s_scale[4 * q], sf = *(const float4*)&s_shift[4 * q];
      const bool pair = out_lo != nullptr;
#pragma unroll 4
      for (int r = rl; r < m; r += 32) {
        const long long i = ((long long)gi * m + r) * c + ch4;
        float4 v = *(const float4*)(x + i);
        v.x = fmaf(v.x, sc.x, sf.x); v.y = fmaf(v.y, sc.y, sf.y);
        v.z = fmaf(v.z, sc.z, sf.z); v.w = fmaf(v.w, sc.w, sf.w);
        if (residual) {
          const float4 rr = *(const float4*)(residual + i);
          v.x += rr.x; v.y += rr.y; v.z += rr.z; v.w += rr.w;
        }
        if (relu) {
          v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
        }
        if (out) *(float4*)(out + i) = v;
        if (out_hi) {
          plane_t h0, l0, h1, l1, h2, l2, h3, l3;
          split16(v.x, pair, h0, l0); split16(v.y, pair, h1, l1);
          split16(v.z, pair, h2, l2); split16(v.w, pair, h3, l3);
          *(uint2*)(out_hi + i) = make_uint2(pack16x2(h0, h1), pack16x2(h2, h3));
          if (pair) *(uint2*)(out_lo + i) = make_uint2(pack16x2(l0, l1), pack16x2(l2, l3));
        }
      }
    }
    __syncthreads();          // s_scale / s_sum are reused by the next grid
  }
}

extern "C" int drb_bn_small_supported(long long m, int c) { return m > 0 && m <= kBnSmallMaxRows && c > 0 && c % 4 == 0; }

extern "C" int drb_bn_small(const float* x, int g, long long m, int c, const float* gamma, const float* beta,
                            float* running_mean, float* running_var, int training, float momentum, float eps,
                            const float* residual, int relu, float* out, void* out_hi, void* out_lo, float* save_mean,
                            float* save_rstd, float* save_scale, float* save_shift, cudaStream_t stream) {
  DRB_REQUIRE(x && g > 0 && (out || out_hi), "drb_bn_small: bad arguments");
  DRB_REQUIRE(drb_bn_small_supported(m, c), "drb_bn_small: m = %lld rows per grid / c = %d not supported (m <= %d, c %% 4 == 0)",
              m, c, kBnSmallMaxRows);
  DRB_REQUIRE(training || (running_mean && running_var), "drb_bn_small: missing statistics");
  DRB_REQUIRE((save_mean == nullptr) == (save_rstd == nullptr) && (save_mean == nullptr) == (save_scale == nullptr) &&
                  (save_mean == nullptr) == (save_shift == nullptr), "drb_bn_small: save buffers come as a set");
  bn_small_kernel<<<(c + 31) / 32, 256, 0, stream>>>(x, g, (int)m, c, gamma, beta, running_mean, running_var, training,
                                                     momentum, eps, residual, relu, out, (plane_t*)out_hi,
                                                     (plane_t*)out_lo, save_mean, save_rstd, save_scale, save_shift);
  DRB_LAUNCH_OK();
  return 0;
}

// ------------------------------------------------------------------------------------------
__global__ void maxpool_kernel(const float* __restrict__ x, int g, int d, int h, int w, int c,
                               int od, int oh, int ow, float* __restrict__ out,
                               plane_t* __restrict__ out_hi, plane_t* __restrict__ out_lo) {
  const long long total = (long long)g * od * oh * ow * c;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int ch = (int)(i % c);
    long long r = i / c;
    const int ox = (int)(r % ow); r /= ow;
    const int oy = (int)(r % oh); r /= oh;
    const int oz = (int)(r % od); r /= od;
    const int gi = (int)r;
    float best = -INFINITY;
    for (int kz = 0; kz < 3; ++kz) {
      const int iz = oz * 2 - 1 + kz;
      if (iz < 0 || iz >= d) continue;
      for (int ky = 0; ky < 3; ++ky) {
        const int iy = oy * 2 - 1 + ky;
        if (iy < 0 || iy >= h) continue;
        for (int kx = 0; kx < 3; ++kx) {
          const int ix = ox * 2 - 1 + kx;
          if (ix < 0 || ix >= w) continue;
          best = fmaxf(best, x[((((long long)gi * d + iz) * h + iy) * w + ix) * c + ch]);
        }
      }
    }
    if (out) out[i] = best;
    if (out_hi) {
      plane_t hh, ll;
      split16(best, out_lo != nullptr, hh, ll);
      out_hi[i] = hh;
      if (out_lo) out_lo[i] = ll;
    }
  }
}

// four channels per thread: 16-byte loads, a quarter of the index arithmetic (fmaxf is exact: same result)
__global__ void __launch_bounds__(256) maxpool4_kernel(const float* __restrict__ x, int g, int d, int h, int w, int c,
                                                       int od, int oh, int ow, float* __restrict__ out,
                                                       plane_t* __restrict__ out_hi, plane_t* __restrict__ out_lo) {
  const int c4n = c >> 2;
  const long long total = (long long)g * od * oh * ow * c4n;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const bool pair = out_lo != nullptr;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int ch = (int)(i % c4n) * 4;
    int r = (int)(i / c4n);                      // output voxels < 2^31 (checked by the host)
    const int ox = r % ow; r /= ow;
    const int oy = r % oh; r /= oh;
    const int oz = r % od; r /= od;
    const int gi = r;
    float4 best = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
    for (int kz = 0; kz < 3; ++kz) {
      const int iz = oz * 2 - 1 + kz;
      if (iz < 0 || iz >= d) continue;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int iy = oy * 2 - 1 + ky;
        if (iy < 0 || iy >= h) continue;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int ix = ox * 2 - 1 + kx;
          if (ix < 0 || ix >= w) continue;
          const float4 v = *(const float4*)(x + ((((long long)gi * d + iz) * h + iy) * w + ix) * c + ch);
          best.x = fmaxf(best.x, v.x); best.y = fmaxf(best.y, v.y);
          best.z = fmaxf(best.z, v.z); best.w = fmaxf(best.w, v.w);
        }
      }
    }
    if (out) *(float4*)(out + i * 4) = best;
    if (out_hi) {
      plane_t h0, l0, h1, l1, h2, l2, h3, l3;
      split16(best.x, pair, h0, l0); split16(best.y, pair, h1, l1);
      split16(best.z, pair, h2, l2); split16(best.w, pair, h3, l3);
      *(uint2*)(out_hi + i * 4) = make_uint2(pack16x2(h0, h1), pack16x2(h2, h3));
      if (pair) *(uint2*)(out_lo + i * 4) = make_uint2(pack16x2(l0, l1), pack16x2(l2, l3));
    }
  }
}

extern "C" int drb_maxpool3d(const float* x, int g, int d, int h, int w, int c, float* out,
                             void* out_hi, void* out_lo, cudaStream_t stream) {
  DRB_REQUIRE(x && (out || out_hi) && g > 0 && d > 0 && h > 0 && w > 0 && c > 0,
              "drb_maxpool3d: bad arguments");
  const int od = (d - 1) / 2 + 1, oh = (h - 1) / 2 + 1, ow = (w - 1) / 2 + 1;
  const long long total = (long long)g * od * oh * ow * c;
  const bool al = ((uintptr_t)x & 15) == 0 && (!out || ((uintptr_t)out & 15) == 0) && (!out_hi || ((uintptr_t)out_hi & 7) == 0) &&
                  (!out_lo || ((uintptr_t)out_lo & 7) == 0);
  if (c % 4 == 0 && al && (long long)g * od * oh * ow < (1LL << 31)) {
    maxpool4_kernel<<<grid_for(total / 4, 256, 148 * 32), 256, 0, stream>>>(
        x, g, d, h, w, c, od, oh, ow, out, (plane_t*)out_hi, (plane_t*)out_lo);
    DRB_LAUNCH_OK();
    return 0;
  }
  maxpool_kernel<<<grid_for(total, 256, 148 * 32), 256, 0, stream>>>(
      x, g, d, h, w, c, od, oh, ow, out, (plane_t*)out_hi, (plane_t*)out_lo);
  DRB_LAUNCH_OK();
  return 0;
}

// ------------------------------------------------------------------------------------------
__global__ void upsample_add_kernel(const float* __restrict__ coarse, int dc, int hc, int wc,
                                    const float* __restrict__ lateral, int g, int d, int h, int w,
                                    int c, float* __restrict__ out, plane_t* __restrict__ out_hi,
                                    plane_t* __restrict__ out_lo) {
  const int c4 = c >> 2;
  const long long total4 = (long long)g * d * h * w * c4;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += stride) {
    const int cc = (int)(i % c4) * 4;
    long long r = i / c4;
    const int x = (int)(r % w); r /= w;
    const int y = (int)(r % h); r /= h;
    const int z = (int)(r % d); r /= d;
    const int gi = (int)r;
    const float4 a = *(const float4*)(lateral + i * 4);
    const float4 b = *(const float4*)(
        coarse + ((((long long)gi * dc + (z >> 1)) * hc + (y >> 1)) * wc + (x >> 1)) * c + cc);
    // torch.add(upsampled, lateral): upsampled is the first operand (feature_pyramid_net.py:74)
    float4 v = make_float4(b.x + a.x, b.y + a.y, b.z + a.z, b.w + a.w);
    if (out) *(float4*)(out + i * 4) = v;
    if (out_hi) {
      const bool pair = out_lo != nullptr;
      plane_t h0, l0, h1, l1, h2, l2, h3, l3;
      split16(v.x, pair, h0, l0); split16(v.y, pair, h1, l1);
      split16(v.z, pair, h2, l2); split16(v.w, pair, h3, l3);
      *(uint2*)(out_hi + i * 4) = make_uint2(pack16x2(h0, h1), pack16x2(h2, h3));
      if (pair) *(uint2*)(out_lo + i * 4) = make_uint2(pack16x2(l0, l1), pack16x2(l2, l3));
    }
  }
}

extern "C" int drb_upsample2_add(const float* coarse, int dc, int hc, int wc, const float* lateral,
                                 int g, int d, int h, int w, int c, float* out, void* out_hi,
                                 void* out_lo, cudaStream_t stream) {
  DRB_REQUIRE(coarse && lateral && (out || out_hi) && c % 4 == 0, "drb_upsample2_add: bad arguments");
  DRB_REQUIRE(d <= 2 * dc && h <= 2 * hc && w <= 2 * wc, "drb_upsample2_add: lateral larger than 2x coarse");
  const long long total4 = (long long)g * d * h * w * (c / 4);
  upsample_add_kernel<<<grid_for(total4, 256, 148 * 32), 256, 0, stream>>>(
      coarse, dc, hc, wc, lateral, g, d, h, w, c, out, (plane_t*)out_hi, (plane_t*)out_lo);
  DRB_LAUNCH_OK();
  return 0;
}

}  // namespace drb
