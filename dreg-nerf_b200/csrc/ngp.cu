// A1-A6: the extract half.  Instant-NGP field query (multiresolution hash grid, 16 levels x 2
// features, T = 2^19, base 16, per-level scale 1.4472692; MLP 32 -> 64 ReLU -> 16; colour head
// SH-4 (16) + 15 features -> 64 -> 64 -> 3 sigmoid) as ONE fused kernel per query type, the
// nerfacc-0.3.5 style occupancy-grid ray marcher for the surface-field mask, and the
// voxel_grid.pt scatter.  Restated from the published algorithms of tiny-cuda-nn / nerfacc (both
// un-vendored by the reference, so PARITY UNPINNED against those wheels); call sites:
// conerf/radiance_fields/ngp.py:92-193, conerf/register/sample_grid.py:208-343,
// conerf/utils/nerfacc_utils.py:84-222, eval_ngp_nerf.py:337-412.
//
// Memory plan: the two finest-reuse levels (0 and 1, dense 16^3 and 24^3, 140 KiB fp32) are staged
// into shared memory with 1-D bulk TMA (cp.async.bulk) once per CTA; levels 2-15 are served from
// L2 (the whole 48 MB table is L2 resident on B200) with 64-bit gathers; MLP weights live in
// shared memory and are read as warp-uniform broadcasts.
#include "common.cuh"
#include "march_math.h"

#include <cub/cub.cuh>
#include <math.h>
#include <stdlib.h>

#include <vector>

namespace drb {


static constexpr int kLevels = 16;
static constexpr int kHashSize = 1 << 19;
#ifndef DRB_SMEM_LEVELS
#define DRB_SMEM_LEVELS 2
#endif
static constexpr int kSmemLevels = DRB_SMEM_LEVELS;

struct LevelTable {
  float scale[kLevels];
  uint32_t res[kLevels];
  uint32_t size[kLevels];     // entries in level
  uint32_t offset[kLevels];   // first entry
  uint32_t total;
};

static LevelTable host_levels() {
  LevelTable t;
  uint32_t off = 0;
  const double b = 1.4472692012786865;
  for (int l = 0; l < kLevels; ++l) {
    // evaluated in double, rounded once to fp32 so that host libraries cannot disagree by an ulp
    const float scale = (float)(exp2((double)l * log2(b)) * 16.0 - 1.0);
    const uint32_t res = (uint32_t)ceilf(scale) + 1u;
    uint64_t n = (uint64_t)res * res * res;
    n = (n + 7) / 8 * 8;
    if (n > (uint64_t)kHashSize) n = kHashSize;
    t.scale[l] = scale; t.res[l] = res; t.size[l] = (uint32_t)n; t.offset[l] = off;
    off += (uint32_t)n;
  }
  t.total = off;
  return t;
}

struct NgpDev {
  const float2* table;
  const float *w1, *w2, *c1, *c2, *c3;
  float amin[3], ainv[3];   // aabb min, 1 / extent
  LevelTable lv;
  // marcher only: cell-major copy of the dense levels 1..4 - the 8 corners of a cell in one 64-byte record
  // (2 sectors per sample and level instead of up to 8); cm_off[l] = first cell of level l
  const float4* cm;
  uint32_t cm_off[5];
};

static NgpDev make_dev(const drb_ngp_params* p) {
  NgpDev d;
  d.table = (const float2*)p->hash_table;
  d.w1 = p->w1; d.w2 = p->w2; d.c1 = p->c1; d.c2 = p->c2; d.c3 = p->c3;
  d.cm = nullptr;
  for (int i = 0; i < 5; ++i) d.cm_off[i] = 0;
  for (int i = 0; i < 3; ++i) {
    d.amin[i] = p->aabb[i];
    d.ainv[i] = p->aabb[3 + i] - p->aabb[i];
  }
  d.lv = host_levels();
  return d;
}

extern "C" long long drb_ngp_table_entries(void) { return (long long)host_levels().total; }

// Shared memory layout of the field kernels.
struct FieldSmem {
  float2* lvl;      // staged levels 0..kSmemLevels-1
  float* w1;        // [64][32]
  float* w2;        // [16][64]
};

__device__ __forceinline__ uint32_t grid_index(const uint32_t g[3], uint32_t res, uint32_t size) {
  uint32_t stride = 1, index = 0;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    if (stride <= size) {
      index += g[d] * stride;
      stride *= res;
    }
  }
  if (size < stride) index = (g[0] * 1u) ^ (g[1] * 2654435761u) ^ (g[2] * 805459861u);
  // NB: replacing this modulo by a mask for the 2^19-entry levels (and a compare for the dense ones) was
  // measured 30 % SLOWER on B200 (same box A/B, round 1), so the plain form stays.
  return index % size;
}

// 32 encoded features of a point already normalised to the unit cube.
__device__ __forceinline__ void hash_encode(const NgpDev& p, const FieldSmem& sm, const float xn[3],
                                            float f[32]) {
#pragma unroll 1
  for (int l = 0; l < kLevels; ++l) {
    const float scale = p.lv.scale[l];
    const uint32_t res = p.lv.res[l], size = p.lv.size[l];
    float frac[3];
    uint32_t g0[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const float pos = fmaf(xn[d], scale, 0.5f);
      const float fl = floorf(pos);
      frac[d] = pos - fl;
      g0[d] = (uint32_t)(int)fl;
    }
    float a0 = 0.f, a1 = 0.f;
    const float2* base = (l < kSmemLevels) ? sm.lvl + p.lv.offset[l] : p.table + p.lv.offset[l];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      uint32_t g[3];
      float w = 1.f;
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        if (c & (1 << d)) { g[d] = g0[d] + 1u; w *= frac[d]; }
        else { g[d] = g0[d]; w *= 1.f - frac[d]; }
      }
      const uint32_t idx = grid_index(g, res, size);
      const float2 v = (l < kSmemLevels) ? base[idx] : __ldg(base + idx);
      a0 = fmaf(w, v.x, a0);
      a1 = fmaf(w, v.y, a1);
    }
    f[2 * l] = a0;
    f[2 * l + 1] = a1;
  }
}

// MLP 32 -> 64 (ReLU) -> NOUT (linear); weights in shared memory (warp-uniform reads).
template <int NOUT>
__device__ __forceinline__ void density_mlp(const FieldSmem& sm, const float f[32], float out[NOUT]) {
#pragma unroll
  for (int o = 0; o < NOUT; ++o) out[o] = 0.f;
#pragma unroll 4
  for (int j = 0; j < 64; ++j) {
    float h = 0.f;
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float4 w = *(const float4*)(sm.w1 + j * 32 + i);
      h = fmaf(w.x, f[i], h); h = fmaf(w.y, f[i + 1], h);
      h = fmaf(w.z, f[i + 2], h); h = fmaf(w.w, f[i + 3], h);
    }
    h = fmaxf(h, 0.f);
#pragma unroll
    for (int o = 0; o < NOUT; ++o) out[o] = fmaf(sm.w2[o * 64 + j], h, out[o]);
  }
}

// Stages levels 0..1 (bulk TMA) and the density MLP weights into shared memory.
__device__ __forceinline__ FieldSmem stage_field(const NgpDev& p, uint8_t* smem, uint64_t* bar) {
  FieldSmem sm;
  const uint32_t lvl_entries = p.lv.offset[kSmemLevels];
  sm.lvl = (float2*)smem;
  sm.w1 = (float*)(smem + (size_t)lvl_entries * sizeof(float2));
  sm.w2 = sm.w1 + 64 * 32;
  const uint32_t bytes = lvl_entries * (uint32_t)sizeof(float2);
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(bar), 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(smem_u32(bar), bytes);
    // 1-D bulk copies are limited in size per instruction; issue in 32 KiB pieces
    for (uint32_t off = 0; off < bytes; off += 32768u) {
      const uint32_t n = bytes - off < 32768u ? bytes - off : 32768u;
      asm volatile(
          "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
          ::"r"(smem_u32(smem + off)), "l"((uint64_t)((const uint8_t*)p.table + off)), "r"(n),
            "r"(smem_u32(bar))
          : "memory");
    }
  }
  for (int i = threadIdx.x; i < 64 * 32; i += blockDim.x) sm.w1[i] = p.w1[i];
  for (int i = threadIdx.x; i < 16 * 64; i += blockDim.x) sm.w2[i] = p.w2[i];
  mbar_wait(smem_u32(bar), 0, nullptr, 0);
  __syncthreads();
  return sm;
}

static size_t field_smem_bytes(const LevelTable& lv) {
  return (size_t)lv.offset[kSmemLevels] * sizeof(float2) + (64 * 32 + 16 * 64) * sizeof(float) + 16;
}

__device__ __forceinline__ bool normalise(const NgpDev& p, const float x[3], float xn[3]) {
  bool inside = true;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    xn[d] = __fdiv_rn(x[d] - p.amin[d], p.ainv[d]);
    inside = inside && (xn[d] > 0.f) && (xn[d] < 1.f);
  }
  return inside;
}

// ------------------------------------------------------------------------------------------
// A1: density (+ 15 geometry features)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512, 1)
ngp_density_kernel(const NgpDev p, const float* __restrict__ x, int n, float* __restrict__ density,
                   float* __restrict__ feat) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  const FieldSmem sm = stage_field(p, smem, &bar);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float xw[3] = {x[i * 3], x[i * 3 + 1], x[i * 3 + 2]};
    float xn[3];
    const bool inside = normalise(p, xw, xn);
    float f[32], o[16];
    hash_encode(p, sm, xn, f);
    density_mlp<16>(sm, f, o);
    density[i] = inside ? expf(o[0] - 1.f) : 0.f;
    if (feat) {
#pragma unroll
      for (int k = 0; k < 15; ++k) feat[(long long)i * 15 + k] = o[1 + k];
    }
  }
}

extern "C" int drb_ngp_density(const drb_ngp_params* pp, const float* x, int n, float* density,
                               float* feat, cudaStream_t stream) {
  DRB_REQUIRE(pp && pp->hash_table && pp->w1 && pp->w2 && x && density, "drb_ngp_density: null argument");
  if (n == 0) return 0;
  const NgpDev p = make_dev(pp);
  const size_t smem = field_smem_bytes(p.lv);
  // set on every call: the attribute is per device and a process may drive several GPUs
  DRB_CUDA_OK(cudaFuncSetAttribute(ngp_density_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = cdiv(n, 512);
  if (grid > 148) grid = 148;
  ngp_density_kernel<<<grid, 512, smem, stream>>>(p, x, n, density, feat);
  DRB_LAUNCH_OK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// A2: colour head, mean over the fixed view directions.  The SH part of the first layer does not
// depend on the point, so per direction it is folded into a 64-vector once per CTA.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void sh4(float x, float y, float z, float o[16]) {
  const float xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
  o[0] = 0.28209479177387814f;
  o[1] = -0.48860251190291987f * y;
  o[2] = 0.48860251190291987f * z;
  o[3] = -0.48860251190291987f * x;
  o[4] = 1.0925484305920792f * xy;
  o[5] = -1.0925484305920792f * yz;
  o[6] = 0.94617469575755997f * z2 - 0.31539156525251999f;
  o[7] = -1.0925484305920792f * xz;
  o[8] = 0.54627421529603959f * x2 - 0.54627421529603959f * y2;
  o[9] = 0.59004358992664352f * y * (-3.0f * x2 + y2);
  o[10] = 2.8906114426405538f * xy * z;
  o[11] = 0.45704579946446572f * y * (1.0f - 5.0f * z2);
  o[12] = 0.3731763325901154f * z * (5.0f * z2 - 3.0f);
  o[13] = 0.45704579946446572f * x * (1.0f - 5.0f * z2);
  o[14] = 1.4453057213202769f * z * (x2 - y2);
  o[15] = 0.59004358992664352f * x * (-x2 + 3.0f * y2);
}

static constexpr int kMaxDirs = 32;
struct DirTable { float d[kMaxDirs][3]; int n; };

// Tensor-core version: one warp owns 32 points (two m16 tiles).  Layer 1's feature part
// (15 -> 64) and layer 2 (64 -> 64, once per direction) run as 3xTF32 mma.sync m16n8k8
// (hi*hi + lo*hi + hi*lo, fp32 accumulate, ~2^-21 relative).  The accumulator layout of one layer
// is fed straight back as the A operand of the next one: within every block of 8 hidden units the
// K index is permuted (fragment column t <-> unit 2t, t+4 <-> 2t+1), and the B fragments are
// pre-arranged with the same permutation, so no shuffle or shared-memory round trip is needed.
// Layer 3 (64 -> 3) + sigmoid + mean over the directions happen in the accumulator layout with one
// quad reduction per direction.
static constexpr int kRgbThreads = 256;
struct RgbSmem {
  float4 b2[8 * 8 * 32];      // layer 2 B fragments [ks][nt][lane] = (hi b0, hi b1, lo b0, lo b1)
  float4 b1[2 * 8 * 32];      // layer 1 (feature columns 16..30 of c1, column 31 handled in dir[])
  float dir[kMaxDirs][64];    // SH (+ padded-one column) contribution of layer 1 per direction
  float c3[3][64];
};

__device__ __forceinline__ uint32_t f2u(float x) { return __float_as_uint(x); }
__device__ __forceinline__ float tf32_trunc(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }
__device__ __forceinline__ void mma_1688(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// acc (+)= A(hi, lo) x B(hi, lo) for one k-step / n-tile, small terms first
__device__ __forceinline__ void mma3(float (&c)[4], const uint32_t (&ahi)[4], const uint32_t (&alo)[4], const float4& b) {
  mma_1688(c, alo, f2u(b.x), f2u(b.y));
  mma_1688(c, ahi, f2u(b.z), f2u(b.w));
  mma_1688(c, ahi, f2u(b.x), f2u(b.y));
}

__global__ void __launch_bounds__(kRgbThreads, 1)
ngp_rgb_kernel(const NgpDev p, const float* __restrict__ feat, int n_all, const DirTable dirs,
               float* __restrict__ rgb, const int* __restrict__ idx, const int* __restrict__ idx_count) {
  // optional row list: evaluate only the points idx[0 .. *idx_count) (rows are independent: same bits per row)
  const int n = idx_count ? *idx_count : n_all;
  extern __shared__ __align__(16) uint8_t rgb_smem_raw[];
  RgbSmem& sm = *(RgbSmem*)rgb_smem_raw;
  // ---- staging: B fragments with the permuted K order, per-direction SH contributions ----
  for (int i = threadIdx.x; i < 8 * 8 * 32; i += blockDim.x) {
    const int ln = i & 31, nt = (i >> 5) & 7, ks = i >> 8;
    const int nn = 8 * nt + (ln >> 2), k0 = 8 * ks + 2 * (ln & 3);
    const float w0 = p.c2[nn * 64 + k0], w1 = p.c2[nn * 64 + k0 + 1];
    const float h0 = tf32_trunc(w0), h1 = tf32_trunc(w1);
    sm.b2[i] = make_float4(h0, h1, tf32_trunc(w0 - h0), tf32_trunc(w1 - h1));
  }
  for (int i = threadIdx.x; i < 2 * 8 * 32; i += blockDim.x) {
    const int ln = i & 31, nt = (i >> 5) & 7, ks = i >> 8;
    const int nn = 8 * nt + (ln >> 2), k0 = 8 * ks + 2 * (ln & 3);
    // feature q sits in column 16 + q of c1; q = 15 is the padding column (its input is the constant 1,
    // folded into dir[] below), so it contributes nothing here
    const float w0 = p.c1[nn * 32 + 16 + k0], w1 = (k0 + 1 < 15) ? p.c1[nn * 32 + 16 + k0 + 1] : 0.f;
    const float h0 = tf32_trunc(w0), h1 = tf32_trunc(w1);
    sm.b1[i] = make_float4(h0, h1, tf32_trunc(w0 - h0), tf32_trunc(w1 - h1));
  }
  for (int i = threadIdx.x; i < 3 * 64; i += blockDim.x) sm.c3[i / 64][i % 64] = p.c3[i];
  for (int i = threadIdx.x; i < dirs.n * 64; i += blockDim.x) {
    const int k = i / 64, j = i % 64;
    // (dir + 1) / 2 -> tcnn maps back to [-1, 1] before evaluating the basis (ngp.py:181)
    const float dx = ((dirs.d[k][0] + 1.f) * 0.5f) * 2.f - 1.f;
    const float dy = ((dirs.d[k][1] + 1.f) * 0.5f) * 2.f - 1.f;
    const float dz = ((dirs.d[k][2] + 1.f) * 0.5f) * 2.f - 1.f;
    float sh[16];
    sh4(dx, dy, dz, sh);
    float acc = p.c1[j * 32 + 31];           // width padding column is fed with 1 (tcnn Identity pad)
#pragma unroll
    for (int q = 0; q < 16; ++q) acc = fmaf(p.c1[j * 32 + q], sh[q], acc);
    sm.dir[k][j] = acc;
  }
  __syncthreads();

  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int warps_total = gridDim.x * (blockDim.x >> 5);
  const int n_batches = (n + 31) / 32;
  for (int batch = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); batch < n_batches; batch += warps_total) {
    const int row0 = batch * 32;
    // ---- layer 1, feature part: h1[32 x 64] = E[32 x 16] * C1f^T, accumulator layout ----
    float h1[2][8][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) h1[mt][nt][i] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      uint32_t ahi[2][4], alo[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          // fragment (row g + 8*(i&1), column t + 4*(i>>1)) <-> feature 8*ks + 2t + (i>>1)
          const int r = row0 + 16 * mt + g + 8 * (i & 1), q = 8 * ks + 2 * t + (i >> 1);
          float v = 0.f;
          if (r < n && q < 15) v = feat[(long long)(idx ? idx[r] : r) * 15 + q];
          const float hi = tf32_trunc(v);
          ahi[mt][i] = f2u(hi);
          alo[mt][i] = f2u(tf32_trunc(v - hi));
        }
      }
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const float4 b = sm.b1[(ks * 8 + nt) * 32 + lane];
        mma3(h1[0][nt], ahi[0], alo[0], b);
        mma3(h1[1][nt], ahi[1], alo[1], b);
      }
    }
    // ---- per direction: layer 2 on the tensor cores, layer 3 + sigmoid in the accumulator layout ----
    float sum[3] = {0.f, 0.f, 0.f};          // lane (g, t) accumulates the colour of row g + 8t
    for (int k = 0; k < dirs.n; ++k) {
      float acc[2][8][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[mt][nt][i] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        // A fragment of k-step ks = relu(h1 + dir) of hidden units 8*ks + {2t, 2t+1}: exactly what this lane
        // holds in h1[mt][ks][*] (accumulator (g, 2t), (g, 2t+1), (g+8, 2t), (g+8, 2t+1))
        const float2 dv = *(const float2*)&sm.dir[k][8 * ks + 2 * t];
        uint32_t ahi[2][4], alo[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          const float v[4] = {fmaxf(h1[mt][ks][0] + dv.x, 0.f), fmaxf(h1[mt][ks][2] + dv.x, 0.f),
                              fmaxf(h1[mt][ks][1] + dv.y, 0.f), fmaxf(h1[mt][ks][3] + dv.y, 0.f)};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float hi = tf32_trunc(v[i]);
            ahi[mt][i] = f2u(hi);
            alo[mt][i] = f2u(tf32_trunc(v[i] - hi));
          }
        }
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          const float4 b = sm.b2[(ks * 8 + nt) * 32 + lane];
          mma3(acc[0][nt], ahi[0], alo[0], b);
          mma3(acc[1][nt], ahi[1], alo[1], b);
        }
      }
      // layer 3: o[c] = sum_j2 c3[c][j2] * relu(h2[j2]); this lane holds units 8*nt + {2t, 2t+1}
      float o[4][3];                           // rows g, g+8, g+16, g+24
#pragma unroll
      for (int r = 0; r < 4; ++r) { o[r][0] = 0.f; o[r][1] = 0.f; o[r][2] = 0.f; }
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float2 w = *(const float2*)&sm.c3[c][8 * nt + 2 * t];
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            o[2 * mt][c] = fmaf(w.x, fmaxf(acc[mt][nt][0], 0.f), o[2 * mt][c]);
            o[2 * mt][c] = fmaf(w.y, fmaxf(acc[mt][nt][1], 0.f), o[2 * mt][c]);
            o[2 * mt + 1][c] = fmaf(w.x, fmaxf(acc[mt][nt][2], 0.f), o[2 * mt + 1][c]);
            o[2 * mt + 1][c] = fmaf(w.y, fmaxf(acc[mt][nt][3], 0.f), o[2 * mt + 1][c]);
          }
        }
      }
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          o[r][c] += __shfl_xor_sync(0xffffffffu, o[r][c], 1);
          o[r][c] += __shfl_xor_sync(0xffffffffu, o[r][c], 2);
        }
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float mine = t == 0 ? o[0][c] : (t == 1 ? o[1][c] : (t == 2 ? o[2][c] : o[3][c]));
        sum[c] += 1.f / (1.f + expf(-mine));
      }
    }
    const int r = row0 + g + 8 * t;
    if (r < n) {
      const float inv = 1.f / (float)dirs.n;
      const long long o3 = (long long)(idx ? idx[r] : r) * 3;
      rgb[o3] = sum[0] * inv; rgb[o3 + 1] = sum[1] * inv; rgb[o3 + 2] = sum[2] * inv;
    }
  }
}

// idx / idx_count (device, optional): evaluate only the listed points (drb_extract_block: the cells that passed both masks)
static int rgb_mean_impl(const drb_ngp_params* pp, const float* feat, int n, const float* host_dirs, int ndirs,
                         float* rgb, const int* idx, const int* idx_count, cudaStream_t stream);
extern "C" int drb_ngp_rgb_mean(const drb_ngp_params* pp, const float* feat, int n, const float* host_dirs,
                                int ndirs, float* rgb, cudaStream_t stream) {
  return rgb_mean_impl(pp, feat, n, host_dirs, ndirs, rgb, nullptr, nullptr, stream);
}
static int rgb_mean_impl(const drb_ngp_params* pp, const float* feat, int n, const float* host_dirs, int ndirs,
                         float* rgb, const int* idx, const int* idx_count, cudaStream_t stream) {
  DRB_REQUIRE(pp && pp->c1 && pp->c2 && pp->c3 && feat && host_dirs && rgb, "drb_ngp_rgb_mean: null argument");
  DRB_REQUIRE(ndirs > 0 && ndirs <= kMaxDirs, "drb_ngp_rgb_mean: 1..%d directions", kMaxDirs);
  if (n == 0) return 0;
  const NgpDev p = make_dev(pp);
  DirTable dt;
  dt.n = ndirs;
  for (int k = 0; k < ndirs; ++k)
    for (int d = 0; d < 3; ++d) dt.d[k][d] = host_dirs[k * 3 + d];
  DRB_CUDA_OK(cudaFuncSetAttribute(ngp_rgb_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RgbSmem)));
  const int warps = kRgbThreads / 32;
  int grid = cdiv(cdiv(n, 32), warps);
  if (grid > igemm_num_sms()) grid = igemm_num_sms();
  ngp_rgb_kernel<<<grid, kRgbThreads, sizeof(RgbSmem), stream>>>(p, feat, n, dt, rgb, idx, idx_count);
  DRB_LAUNCH_OK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// A5: surface-field mask.  One thread per (camera, point) ray; marching restated from nerfacc
// 0.3.5's ray_marching kernel (fixed step, occupancy-grid skipping, AABB contraction).
// ------------------------------------------------------------------------------------------
struct MarchArgs {
  float roi_min[3], roi_max[3], scene_min[3], scene_max[3];
  int res;
  float step, cut_off;
  int max_skips;      // empty-space events a lane may take per outer iteration
  long long watchdog_clocks;   // give up (and raise the device error flag) after this many SM clocks
  uint32_t watchdog_mask;      // rounds between clock reads - 1
  int* err;                    // per-device error flag (code 31 = marcher watchdog)
};

// ------------------------------------------------------------------------------------------
// Compile-time level table (must equal host_levels(); checked once at run time by levels_ok()).
// ------------------------------------------------------------------------------------------
__host__ __device__ constexpr uint32_t lvl_res(int l) {
  constexpr uint32_t r[kLevels] = {16, 24, 34, 49, 71, 102, 148, 213, 308, 446, 646, 934, 1352, 1956, 2831, 4096};
  return r[l];
}
__host__ __device__ constexpr uint32_t lvl_size(int l) {
  constexpr uint32_t s[kLevels] = {4096, 13824, 39304, 117656, 357912, 524288, 524288, 524288,
                                   524288, 524288, 524288, 524288, 524288, 524288, 524288, 524288};
  return s[l];
}
__host__ __device__ constexpr uint32_t lvl_off(int l) {
  uint32_t o = 0;
  for (int i = 0; i < l; ++i) o += lvl_size(i);
  return o;
}
__host__ __device__ constexpr bool lvl_dense(int l) {
  return (uint64_t)lvl_res(l) * lvl_res(l) * lvl_res(l) <= (uint64_t)lvl_size(l);
}
static bool levels_ok(const LevelTable& t) {
  bool ok = true;
  for (int l = 0; l < kLevels; ++l)
    ok = ok && t.res[l] == lvl_res(l) && t.size[l] == lvl_size(l) && t.offset[l] == lvl_off(l);
  return ok;
}

// One level of the hash encoding with every level constant folded: dense levels index
// x + y*res + z*res^2 with a conditional wrap (the index stays below 2 * size for points inside the
// unit cube), hashed levels share the per-axis products between the 8 corners and mask with 2^19 - 1.
// Same arithmetic and summation order as hash_encode(): identical features for 0 < xn < 1.
template <int L, int SMEM_LEVELS>
__device__ __forceinline__ void encode_level(const NgpDev& p, const float2* __restrict__ s_lvl,
                                             const float xn[3], float& o0, float& o1) {
  constexpr uint32_t res = lvl_res(L), size = lvl_size(L), off = lvl_off(L);
  const float scale = p.lv.scale[L];
  float fr[3];
  uint32_t g[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const float pos = fmaf(xn[d], scale, 0.5f);
    const float fl = floorf(pos);
    fr[d] = pos - fl;
    g[d] = (uint32_t)(int)fl;
  }
  // x-neighbours (corners c and c^1) are adjacent table entries whenever the x index of the lower
  // corner is even (dense: idx, idx + 1; hashed: idx, idx ^ 1), i.e. one aligned 16-byte load serves
  // both.  The load/store unit is bound by wavefronts (one per distinct sector per instruction), so
  // this removes a quarter of them; odd x indices fetch the upper corner separately.  (Reading the fine
  // levels with L1::no_allocate was measured 10 % slower on B200 and is not used.)
  uint32_t idx0[4], idx1[4];              // lower / upper x corner of the 4 (y, z) combinations
  if constexpr (lvl_dense(L)) {
    const uint32_t base = g[0] + g[1] * res + g[2] * (res * res);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const uint32_t i = base + ((c & 1) ? res : 0u) + ((c & 2) ? res * res : 0u);
      idx0[c] = i >= size ? i - size : i;
      idx1[c] = i + 1u >= size ? i + 1u - size : i + 1u;
    }
  } else {
    static_assert(lvl_dense(L) || (size & (size - 1)) == 0, "hashed levels are powers of two");
    const uint32_t hy0 = g[1] * 2654435761u, hz0 = g[2] * 805459861u;
    const uint32_t hy[2] = {hy0, hy0 + 2654435761u};
    const uint32_t hz[2] = {hz0, hz0 + 805459861u};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const uint32_t hyz = hy[c & 1] ^ hz[c >> 1];
      idx0[c] = (g[0] ^ hyz) & (size - 1u);
      idx1[c] = ((g[0] + 1u) ^ hyz) & (size - 1u);
    }
  }
  float2 v[8];
  if constexpr (L < SMEM_LEVELS) {
#pragma unroll
    for (int c = 0; c < 4; ++c) { v[2 * c] = s_lvl[off + idx0[c]]; v[2 * c + 1] = s_lvl[off + idx1[c]]; }
  } else {
    static_assert(off % 2 == 0 && size % 2 == 0, "16-byte pairs need even level offsets");
    const bool paired = (g[0] & 1u) == 0u;          // dense: idx0 parity == g[0] parity only if res is even
    float4 q[4];
    if constexpr (lvl_dense(L) && (res & 1u)) {
      // odd resolution: the parity of idx0 is not that of g[0]; decide per (y, z) combination
#pragma unroll
      for (int c = 0; c < 4; ++c) q[c] = __ldg((const float4*)(p.table + off + (idx0[c] & ~1u)));
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const bool odd = idx0[c] & 1u;
        v[2 * c] = odd ? make_float2(q[c].z, q[c].w) : make_float2(q[c].x, q[c].y);
        if (!odd && idx1[c] == idx0[c] + 1u) v[2 * c + 1] = make_float2(q[c].z, q[c].w);
        else v[2 * c + 1] = __ldg(p.table + off + idx1[c]);
      }
    } else {
#pragma unroll
      for (int c = 0; c < 4; ++c) q[c] = __ldg((const float4*)(p.table + off + (idx0[c] & ~1u)));
      if (paired) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          // dense, even res: idx0 even -> idx1 = idx0 + 1 (no wrap between them: size is even);
          // hashed: idx1 = idx0 ^ 1.  Either way the mate of the aligned pair.
          const bool odd = idx0[c] & 1u;
          v[2 * c] = odd ? make_float2(q[c].z, q[c].w) : make_float2(q[c].x, q[c].y);
          v[2 * c + 1] = odd ? make_float2(q[c].x, q[c].y) : make_float2(q[c].z, q[c].w);
        }
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const bool odd = idx0[c] & 1u;
          v[2 * c] = odd ? make_float2(q[c].z, q[c].w) : make_float2(q[c].x, q[c].y);
          v[2 * c + 1] = __ldg(p.table + off + idx1[c]);
        }
      }
    }
  }
  float a0 = 0.f, a1 = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    float w = 1.f;
#pragma unroll
    for (int d = 0; d < 3; ++d) w *= (c & (1 << d)) ? fr[d] : 1.f - fr[d];
    a0 = fmaf(w, v[c].x, a0);
    a1 = fmaf(w, v[c].y, a1);
  }
  o0 = a0;
  o1 = a1;
}

// ------------------------------------------------------------------------------------------
// Warp-level density evaluation of 32 samples (one per lane): hash encoding on the CUDA cores, four
// levels (8 features = one K-step) at a time through a 32 x 8 shared-memory tile, layer 1 (32 -> 64)
// on the tensor cores as 3xTF32 (hi*hi + lo*hi + hi*lo, ~2^-21 relative), ReLU + layer 2 (64 -> 1)
// + quad shuffle reduction in the accumulator layout.  Tile row pitch 12 words: conflict free for
// the 128-bit row stores and for the m16n8k8 A-fragment loads; B fragments of W1 (hi, lo) are
// pre-arranged per (k-step, n-tile, lane) at staging time.  The tile is kept this small on purpose:
// every KB of shared memory is a KB less L1 for the table gathers, which bound the kernel.
// ------------------------------------------------------------------------------------------
#ifndef DRB_MARCH_SMEM_LEVELS
#define DRB_MARCH_SMEM_LEVELS 1
#endif
static constexpr int kMarchSmemLevels = DRB_MARCH_SMEM_LEVELS;           // hash levels staged in shared memory by the marcher
static constexpr int kTilePitch = 12;                 // words per sample row of the A tile (one K-step)

__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

struct WarpMlp {
  float* tile;            // [32][kTilePitch] this warp's A tile
  const float4* bfrag;    // [4 k-steps][8 n-tiles][32 lanes] = (hi b0, hi b1, lo b0, lo b1)
  const float* w2;        // [64] output row 0 of layer 2
};

// Hashed level with run-time level index: levels 5..15 differ only in scale and table offset, so one
// copy of this body serves all eleven from a rolled loop.  The unrolled form was 5.1 k SASS instructions
// (80 KB) walked once per batch by 16 warps at different places: 33 % of the stall samples were
// instruction-fetch misses (ncu, round 1).  Same arithmetic and summation order as encode_level<L>.
#ifndef DRB_HASH_UNROLL
#define DRB_HASH_UNROLL 1
#endif
static constexpr int kHashUnroll = DRB_HASH_UNROLL;
__device__ __forceinline__ void encode_hashed(const NgpDev& p, int l, const float xn[3], float& o0, float& o1) {
  constexpr uint32_t size = 1u << 19;
  const float scale = p.lv.scale[l];
  const float2* __restrict__ tab = p.table + p.lv.offset[l];
  float fr[3];
  uint32_t g[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const float pos = fmaf(xn[d], scale, 0.5f);
    const float fl = floorf(pos);
    fr[d] = pos - fl;
    g[d] = (uint32_t)(int)fl;
  }
  const uint32_t hy0 = g[1] * 2654435761u, hz0 = g[2] * 805459861u;
  const uint32_t hy[2] = {hy0, hy0 + 2654435761u};
  const uint32_t hz[2] = {hz0, hz0 + 805459861u};
  uint32_t idx0[4], idx1[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const uint32_t hyz = hy[c & 1] ^ hz[c >> 1];
    idx0[c] = (g[0] ^ hyz) & (size - 1u);
    idx1[c] = ((g[0] + 1u) ^ hyz) & (size - 1u);
  }
  float4 q[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) q[c] = __ldg((const float4*)(tab + (idx0[c] & ~1u)));
  float2 v[8];
  if ((g[0] & 1u) == 0u) {                     // idx1 = idx0 ^ 1: the mate of the aligned pair
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const bool odd = idx0[c] & 1u;
      v[2 * c] = odd ? make_float2(q[c].z, q[c].w) : make_float2(q[c].x, q[c].y);
      v[2 * c + 1] = odd ? make_float2(q[c].x, q[c].y) : make_float2(q[c].z, q[c].w);
    }
  } else {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const bool odd = idx0[c] & 1u;
      v[2 * c] = odd ? make_float2(q[c].z, q[c].w) : make_float2(q[c].x, q[c].y);
      v[2 * c + 1] = __ldg(tab + idx1[c]);
    }
  }
  float a0 = 0.f, a1 = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    float w = 1.f;
#pragma unroll
    for (int d = 0; d < 3; ++d) w *= (c & (1 << d)) ? fr[d] : 1.f - fr[d];
    a0 = fmaf(w, v[c].x, a0);
    a1 = fmaf(w, v[c].y, a1);
  }
  o0 = a0;
  o1 = a1;
}

// Dense level with run-time level index (levels 1..4 of the marcher; level 0 lives in shared memory):
// index x + y*res + z*res^2 with the conditional wrap, aligned 16-byte pair when the lower x corner
// is even and its x-neighbour is the next entry.  Same arithmetic as encode_level<L>.
__device__ __forceinline__ void encode_dense(const NgpDev& p, int l, const float xn[3], float& o0, float& o1) {
  const float scale = p.lv.scale[l];
  const uint32_t res = p.lv.res[l], size = p.lv.size[l];
  const float2* __restrict__ tab = p.table + p.lv.offset[l];
  float fr[3];
  uint32_t g[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const float pos = fmaf(xn[d], scale, 0.5f);
    const float fl = floorf(pos);
    fr[d] = pos - fl;
    g[d] = (uint32_t)(int)fl;
  }
  const uint32_t res2 = res * res;
  const uint32_t base = g[0] + g[1] * res + g[2] * res2;
  float2 v[8];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const uint32_t i = base + ((c & 1) ? res : 0u) + ((c & 2) ? res2 : 0u);
    const uint32_t i0 = i >= size ? i - size : i;
    const uint32_t i1 = i + 1u >= size ? i + 1u - size : i + 1u;
    const float4 q = __ldg((const float4*)(tab + (i0 & ~1u)));
    const bool odd = i0 & 1u;
    v[2 * c] = odd ? make_float2(q.z, q.w) : make_float2(q.x, q.y);
    if (!odd && i1 == i0 + 1u) v[2 * c + 1] = make_float2(q.z, q.w);
    else v[2 * c + 1] = __ldg(tab + i1);
  }
  float a0 = 0.f, a1 = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    float w = 1.f;
#pragma unroll
    for (int d = 0; d < 3; ++d) w *= (c & (1 << d)) ? fr[d] : 1.f - fr[d];
    a0 = fmaf(w, v[c].x, a0);
    a1 = fmaf(w, v[c].y, a1);
  }
  o0 = a0;
  o1 = a1;
}

// Cell-major variant of encode_dense: the same 8 table entries (same index arithmetic, copied by
// cellmajor_kernel), read as one 64-byte record -> identical features, 2 sectors instead of up to 8.
__device__ __forceinline__ void encode_dense_cm(const NgpDev& p, int l, const float xn[3], float& o0, float& o1) {
  const float scale = p.lv.scale[l];
  const uint32_t res = p.lv.res[l];
  float fr[3];
  uint32_t g[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const float pos = fmaf(xn[d], scale, 0.5f);
    const float fl = floorf(pos);
    fr[d] = pos - fl;
    g[d] = (uint32_t)(int)fl;
  }
  const uint32_t cell = g[0] + g[1] * res + g[2] * res * res;
  const float4* rec = p.cm + ((size_t)p.cm_off[l] + cell) * 4;
  float2 v[8];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const float4 q = __ldg(rec + c);
    v[2 * c] = make_float2(q.x, q.y);
    v[2 * c + 1] = make_float2(q.z, q.w);
  }
  float a0 = 0.f, a1 = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    float w = 1.f;
#pragma unroll
    for (int d = 0; d < 3; ++d) w *= (c & (1 << d)) ? fr[d] : 1.f - fr[d];
    a0 = fmaf(w, v[c].x, a0);
    a1 = fmaf(w, v[c].y, a1);
  }
  o0 = a0;
  o1 = a1;
}

// record[cell][c] = table entry of corner c (bit 0 = x, bit 1 = y, bit 2 = z) of cell (g0, g1, g2), with
// encode_dense's conditional wrap.  One thread per (cell, corner).
__global__ void cellmajor_kernel(const NgpDev p, float2* __restrict__ cm) {
  const uint32_t total = p.cm_off[4] + p.lv.res[4] * p.lv.res[4] * p.lv.res[4];
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total * 8u) return;
  const uint32_t c = i & 7u, cell_g = i >> 3;
  int l = 1;
  while (l < 4 && cell_g >= p.cm_off[l + 1]) ++l;
  const uint32_t cell = cell_g - p.cm_off[l];
  const uint32_t res = p.lv.res[l], size = p.lv.size[l];
  const uint32_t g0 = cell % res, g1 = (cell / res) % res, g2 = cell / (res * res);
  uint32_t idx = g0 + (c & 1u) + (g1 + ((c >> 1) & 1u)) * res + (g2 + ((c >> 2) & 1u)) * res * res;
  if (idx >= size) idx -= size;
  cm[i] = p.table[p.lv.offset[l] + idx];
}

// One K-step of layer 1: the 32 x 8 tile (this warp's rows are complete) -> A fragments -> 48 MMAs.
__device__ __forceinline__ void mlp_kstep(const WarpMlp& m, int ks, float (&acc)[2][8][4], int lane) {
  const int g = lane >> 2, t = lane & 3;
  __syncwarp();                                        // every lane's features of this K-step are in the tile
  uint32_t ahi[2][4], alo[2][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
    const float* r0 = m.tile + (16 * mt + g) * kTilePitch + t;
    const float v[4] = {r0[0], r0[8 * kTilePitch], r0[4], r0[8 * kTilePitch + 4]};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float hi = tf32_hi(v[i]);
      ahi[mt][i] = __float_as_uint(hi);
      alo[mt][i] = __float_as_uint(tf32_hi(v[i] - hi));
    }
  }
  __syncwarp();                                        // fragment loads done: the tile may be overwritten
  const float4* bf = m.bfrag + ks * 8 * 32 + lane;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    const float4 b = bf[nt * 32];
    const uint32_t bh0 = __float_as_uint(b.x), bh1 = __float_as_uint(b.y);
    const uint32_t bl0 = __float_as_uint(b.z), bl1 = __float_as_uint(b.w);
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      mma_tf32(acc[mt][nt], alo[mt], bh0, bh1);
      mma_tf32(acc[mt][nt], ahi[mt], bl0, bl1);
      mma_tf32(acc[mt][nt], ahi[mt], bh0, bh1);
    }
  }
}

// Returns layer-2 output 0 (pre-activation of the density) of the calling lane's sample.
__device__ __forceinline__ float warp_density_raw(const NgpDev& p, const float2* __restrict__ s_lvl,
                                                  const WarpMlp& m, const float xn[3], int lane) {
  float acc[2][8][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[mt][nt][i] = 0.f;
  float* row = m.tile + lane * kTilePitch;
  static_assert(kMarchSmemLevels <= 1, "levels 1.. are read through L1 by the rolled loops");
  {
    float f0, f1;
    encode_level<0, kMarchSmemLevels>(p, s_lvl, xn, f0, f1);       // level 0: shared memory (or L1)
    *(float2*)row = make_float2(f0, f1);
  }
#pragma unroll 1
  for (int l = 1; l < 5; ++l) {
    float f0, f1;
    if (p.cm) encode_dense_cm(p, l, xn, f0, f1); else encode_dense(p, l, xn, f0, f1);
    *(float2*)(row + 2 * (l & 3)) = make_float2(f0, f1);
    if ((l & 3) == 3) mlp_kstep(m, l >> 2, acc, lane);
  }
  static_assert(!lvl_dense(5) && lvl_dense(4), "levels 5..15 are the hashed ones");
#pragma unroll kHashUnroll
  for (int l = 5; l < kLevels; ++l) {
    float f0, f1;
    encode_hashed(p, l, xn, f0, f1);
    *(float2*)(row + 2 * (l & 3)) = make_float2(f0, f1);
    if ((l & 3) == 3) mlp_kstep(m, l >> 2, acc, lane);
  }
  // accumulator layout: acc[mt][nt] = {(row g, col 2t), (g, 2t+1), (g+8, 2t), (g+8, 2t+1)} of tile
  // rows 16*mt.., hidden units 8*nt..
  const int t = lane & 3;
  float s[4] = {0.f, 0.f, 0.f, 0.f};                   // rows g, g+8, g+16, g+24
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    const float2 w = *(const float2*)(m.w2 + 8 * nt + 2 * t);
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      s[2 * mt] = fmaf(w.x, fmaxf(acc[mt][nt][0], 0.f), s[2 * mt]);
      s[2 * mt] = fmaf(w.y, fmaxf(acc[mt][nt][1], 0.f), s[2 * mt]);
      s[2 * mt + 1] = fmaf(w.x, fmaxf(acc[mt][nt][2], 0.f), s[2 * mt + 1]);
      s[2 * mt + 1] = fmaf(w.y, fmaxf(acc[mt][nt][3], 0.f), s[2 * mt + 1]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    s[i] += __shfl_xor_sync(0xffffffffu, s[i], 1);
    s[i] += __shfl_xor_sync(0xffffffffu, s[i], 2);
  }
  // every lane of quad g now holds rows g + 8j; lane t publishes row g + 8t, lane L fetches row L
  const float mine = t == 0 ? s[0] : (t == 1 ? s[1] : (t == 2 ? s[2] : s[3]));
  return __shfl_sync(0xffffffffu, mine, 4 * (lane & 7) + (lane >> 3));
}

// ------------------------------------------------------------------------------------------
// Surface-field marcher.  One ray per (camera, point); per-ray arithmetic (t0/t1/tm updates, skip
// rule, termination tests) is the scalar oracle's, so masks are identical.
//
// Scheduling.  A ray alternates between cheap, divergent empty-space skipping (phase A) and the
// expensive density sample (phase B).  Every warp keeps up to 64 rays in shared-memory slots.  In
// phase A each lane advances one ray (in registers) until its next sample position lies in an
// occupied cell; the ray is then parked in its slot on the warp's pending list and the lane takes
// another ray.  As soon as 8 rays are pending the warp evaluates them together - up to 4 consecutive
// samples per ray (see "speculation" in phase A), 32 lanes, layer 1 on the tensor cores - and the rays
// go back to the resume list.  New rays come in windows of 32 consecutive (camera, point) pairs with
// the points in Morton order, so the rays a warp holds are spatial neighbours.  Before this rewrite
// the kernel ran with 14 of 32 lanes active on average (ncu, round 1).
//
// Shared-memory diet.  Phase B is bound by the rate at which an SM's L1 retires distinct 32-byte
// sectors (one per load instruction per clock, halved when the carve-out leaves ~28 KB of L1:
// scripts/ubench/gather*.cu), so L1 capacity matters more than anything kept in shared memory: the plan below (hash level 0, B
// fragments, coarse bitmap, 16 x (A tile + 64 ray slots + lists)) stays under 132 KB, which leaves a
// 96 KB L1 (measured: 29 -> 26 ms on the heaviest synthetic block against the 164 KB carve-out).
//
// Latency.  Phase A is a chain of dependent occupancy lookups.  A coarse "any voxel occupied"
// bitmap (cells of cf^3 voxels, <= 32^3 bits = 4 KB) sits in shared memory and answers most
// lookups of the empty-space approach without leaving the SM; the IEEE divisions by the ROI
// extent are done as q = x*y, r = fma(-e, q, x), q + r*y with y = RN(1/e), which is the correctly
// rounded quotient (Markstein) without the MUFU/slow-path sequence.  Camera origins are staged in
// shared memory as well.
// ------------------------------------------------------------------------------------------
__device__ unsigned long long g_march_stats[4];   // rays, skip events, density samples, warp rounds
extern "C" int drb_march_stats(unsigned long long* host4, int reset) {
  DRB_REQUIRE(host4, "drb_march_stats: null argument");
  DRB_CUDA_OK(cudaMemcpyFromSymbol(host4, g_march_stats, sizeof(unsigned long long) * 4));
  if (reset) {
    const unsigned long long z[4] = {0, 0, 0, 0};
    DRB_CUDA_OK(cudaMemcpyToSymbol(g_march_stats, z, sizeof(z)));
  }
  return 0;
}

#ifndef DRB_MARCH_THREADS
#define DRB_MARCH_THREADS 512
#endif
static constexpr int kMarchThreads = DRB_MARCH_THREADS;
static constexpr int kMarchWarps = kMarchThreads / 32;
static constexpr int kSlots = 64;                    // rays per warp (in flight + pending + resumable)
static constexpr int kSlotWords = 10;                // dir[3] len t0 t1 tm T best (pi | cam << 22)
static constexpr int kPiBits = 22;
#ifndef DRB_SPEC
#define DRB_SPEC 4
#endif
static constexpr int kSpec = DRB_SPEC;               // consecutive samples of one ray evaluated in one batch
static constexpr int kRaysPerBatch = 32 / kSpec;
static_assert(kSpec == 1 || kSpec == 2 || kSpec == 4 || kSpec == 8, "group size");
#ifndef DRB_WINDOW
#define DRB_WINDOW 32
#endif
static constexpr int kWindow = DRB_WINDOW;                  // consecutive rays a warp takes from the global counter per fetch
static constexpr int kCandCap = 32 + kWindow;        // per-warp queue of rays that passed the "not yet seen" test
#ifndef DRB_COARSE_DIM
#define DRB_COARSE_DIM 32
#endif
static constexpr int kCoarseMaxDim = DRB_COARSE_DIM;             // coarse occupancy bitmap: at most DRB_COARSE_DIM^3 bits (32^3 = 4 KB)
static constexpr int kCoarseWords = kCoarseMaxDim * kCoarseMaxDim * kCoarseMaxDim / 32;
static constexpr int kMaxCams = (1 << (32 - kPiBits)) - 1;
static constexpr size_t kWarpBytes = (size_t)32 * kTilePitch * 4 + (size_t)kSlots * kSlotWords * 4 + 4 * kSlots +
                                     (size_t)kCandCap * 4;
static_assert(kMarchSmemLevels <= kSmemLevels, "the marcher stages a prefix of the field kernels' levels");

struct RayState {
  float o[3], dir[3], inv[3];
  float len, t0, t1, tm, T, best;
  int pi, ci;
};

static size_t march_smem_bytes(int ncams) {
  return (size_t)lvl_off(kMarchSmemLevels) * sizeof(float2) + 64 * sizeof(float) + (size_t)4 * 8 * 32 * 16 +
         (size_t)kCoarseWords * 4 + (((size_t)ncams * 12 + 15) & ~(size_t)15) + kMarchWarps * kWarpBytes + 16;
}

// coarse[c] bit = any voxel of the cf^3 cell occupied; cdim = ceil(res / cf) cells per axis.
__global__ void coarse_occ_kernel(const uint8_t* __restrict__ occ, int res, int shift, int cdim,
                                  uint32_t* __restrict__ coarse) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = cdim * cdim * cdim;
  bool any = false;
  if (c < total) {
    const int cz = c % cdim, cy = (c / cdim) % cdim, cx = c / (cdim * cdim);
    const int cf = 1 << shift;
    for (int dx = 0; dx < cf && !any; ++dx)
      for (int dy = 0; dy < cf && !any; ++dy)
        for (int dz = 0; dz < cf; ++dz) {
          const int x = (cx << shift) + dx, y = (cy << shift) + dy, z = (cz << shift) + dz;
          if (x < res && y < res && z < res && occ[((long long)x * res + y) * res + z]) { any = true; break; }
        }
  }
  const uint32_t word = __ballot_sync(0xffffffffu, any);
  if ((threadIdx.x & 31) == 0 && c < ((total + 31) & ~31)) coarse[c >> 5] = word;
}

struct MarchAux {
  const uint32_t* coarse;   // global copy of the coarse bitmap
  int coarse_shift, coarse_dim, coarse_words;
  float roi_rcp[3];         // RN(1 / roi extent)
  float res_rcp;            // RN(1 / res)
};

// correctly rounded a / b given y = RN(1 / b) (no overflow / underflow in this range)
__device__ __forceinline__ float div_rn_rcp(float a, float b, float y) {
  const float q = a * y;
  const float r = fmaf(-b, q, a);
  return fmaf(r, y, q);
}

__global__ void __launch_bounds__(kMarchThreads, 1)
surface_mask_kernel(const NgpDev p, const MarchArgs a, const MarchAux aux, const uint8_t* __restrict__ occ,
                    const float* __restrict__ points, int n, const float* __restrict__ cams, int ncams,
                    const int* __restrict__ active_idx, const int* __restrict__ active_count,
                    unsigned long long* __restrict__ counter, uint8_t* __restrict__ surface) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  // ---- staging: hash level 0 by bulk TMA, W1 as tf32 hi/lo B fragments, w2 row 0, coarse bitmap, cameras
  constexpr uint32_t lvl_bytes = lvl_off(kMarchSmemLevels) * (uint32_t)sizeof(float2);
  float2* s_lvl = (float2*)smem;
  float* s_w2 = (float*)(smem + lvl_bytes);
  float4* s_bfrag = (float4*)(smem + lvl_bytes + 64 * sizeof(float));
  uint32_t* s_coarse = (uint32_t*)(smem + lvl_bytes + 64 * sizeof(float) + (size_t)4 * 8 * 32 * 16);
  float* s_cams = (float*)(s_coarse + kCoarseWords);
  uint8_t* s_warp = (uint8_t*)s_cams + (((size_t)ncams * 12 + 15) & ~(size_t)15);
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (threadIdx.x == 0 && lvl_bytes > 0) {
    mbar_expect_tx(smem_u32(&bar), lvl_bytes);
    for (uint32_t off = 0; off < lvl_bytes; off += 32768u) {
      const uint32_t nb = lvl_bytes - off < 32768u ? lvl_bytes - off : 32768u;
      asm volatile(
          "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
          ::"r"(smem_u32(smem + off)), "l"((uint64_t)((const uint8_t*)p.table + off)), "r"(nb),
            "r"(smem_u32(&bar))
          : "memory");
    }
  }
  for (int i = threadIdx.x; i < 64; i += blockDim.x) s_w2[i] = p.w2[i];
  for (int i = threadIdx.x; i < 4 * 8 * 32; i += blockDim.x) {
    const int ln = i & 31, nt = (i >> 5) & 7, ks = i >> 8;
    const int nn = 8 * nt + (ln >> 2), k0 = 8 * ks + (ln & 3);
    const float b0 = p.w1[nn * 32 + k0], b1 = p.w1[nn * 32 + k0 + 4];
    const float h0 = tf32_hi(b0), h1 = tf32_hi(b1);
    s_bfrag[i] = make_float4(h0, h1, tf32_hi(b0 - h0), tf32_hi(b1 - h1));
  }
  for (int i = threadIdx.x; i < aux.coarse_words; i += blockDim.x) s_coarse[i] = aux.coarse[i];
  for (int i = threadIdx.x; i < ncams * 3; i += blockDim.x) s_cams[i] = cams[i];
  if (lvl_bytes > 0) mbar_wait(smem_u32(&bar), 0, nullptr, 0);
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint8_t* wbase = s_warp + (size_t)warp * kWarpBytes;
  WarpMlp mlp;
  mlp.tile = (float*)wbase;
  mlp.bfrag = s_bfrag;
  mlp.w2 = s_w2;
  float* slots = (float*)(wbase + 32 * kTilePitch * 4);            // [kSlotWords][kSlots]
  uint16_t* l_pend = (uint16_t*)(wbase + 32 * kTilePitch * 4 + kSlots * kSlotWords * 4);   // slot | (nspec-1) << 6
  uint8_t* l_resume = (uint8_t*)(l_pend + kSlots);
  uint8_t* l_free = l_resume + kSlots;
  uint32_t* l_cand = (uint32_t*)(l_free + kSlots);                 // (pi | cam << 22) of rays waiting to start
  for (int i = lane; i < kSlots; i += 32) l_free[i] = (uint8_t)i;
  __syncwarp();
  int n_pend = 0, n_resume = 0, n_free = kSlots, n_cand = 0;       // warp-uniform
  const uint32_t lt_mask = (1u << lane) - 1u;

  const int n_act = active_count ? *active_count : n;
  const unsigned long long total = (unsigned long long)n_act * (unsigned long long)ncams;
  const int kMaxSkips = a.max_skips;
  const float roi_ext[3] = {a.roi_max[0] - a.roi_min[0], a.roi_max[1] - a.roi_min[1], a.roi_max[2] - a.roi_min[2]};

  RayState ray;
  bool have = false;          // this lane is advancing a ray (state in registers, home slot my_slot)
  int my_slot = -1;
  bool global_done = false;   // warp-uniform: the global ray counter is exhausted
  uint32_t st_rays = 0, st_skips = 0, st_samples = 0, st_rounds = 0;   // roofline accounting

  auto store_ray = [&](int s) {
    slots[0 * kSlots + s] = ray.dir[0]; slots[1 * kSlots + s] = ray.dir[1]; slots[2 * kSlots + s] = ray.dir[2];
    slots[3 * kSlots + s] = ray.len; slots[4 * kSlots + s] = ray.t0; slots[5 * kSlots + s] = ray.t1;
    slots[6 * kSlots + s] = ray.tm; slots[7 * kSlots + s] = ray.T; slots[8 * kSlots + s] = ray.best;
    slots[9 * kSlots + s] = __uint_as_float((uint32_t)ray.pi | ((uint32_t)ray.ci << kPiBits));
  };
  // valid == false reads slot 0 / camera 0 into a ray that is never used: the registers are then dead
  // across phase B for every lane (no state has to survive the tensor-core section in registers)
  auto load_ray = [&](int s, bool valid) {
    s = valid ? s : 0;
    ray.dir[0] = slots[0 * kSlots + s]; ray.dir[1] = slots[1 * kSlots + s]; ray.dir[2] = slots[2 * kSlots + s];
    ray.len = slots[3 * kSlots + s]; ray.t0 = slots[4 * kSlots + s]; ray.t1 = slots[5 * kSlots + s];
    ray.tm = slots[6 * kSlots + s]; ray.T = slots[7 * kSlots + s]; ray.best = slots[8 * kSlots + s];
    const uint32_t pc = valid ? __float_as_uint(slots[9 * kSlots + s]) : 0u;
    ray.pi = (int)(pc & ((1u << kPiBits) - 1u)); ray.ci = (int)(pc >> kPiBits);
#pragma unroll
    for (int d = 0; d < 3; ++d) { ray.o[d] = s_cams[ray.ci * 3 + d]; ray.inv[d] = 1.f / ray.dir[d]; }
  };
  const long long t_start = clock64();

  for (uint32_t round = 0;; ++round) {
    // watchdog: a scheduling bug must not hang the GPU box (~30 s at 2 GHz, then give up) - and a truncated
    // march must not pass for a result: the flag makes the host call fail (drb_march_status)
    if ((round & a.watchdog_mask) == a.watchdog_mask && clock64() - t_start > a.watchdog_clocks) {
      if (lane == 0 && a.err) *(volatile int*)a.err = 31;
      break;
    }
    ++st_rounds;
    // ---------------- refill: lanes without a ray take a resumable one, else start new rays ------
    {
      const uint32_t need = __ballot_sync(0xffffffffu, !have);
      // spare slots of lanes whose ray ended in phase A go back to the free list first
      const uint32_t spare = __ballot_sync(0xffffffffu, !have && my_slot >= 0);
      if (!have && my_slot >= 0) { l_free[n_free + __popc(spare & lt_mask)] = (uint8_t)my_slot; my_slot = -1; }
      n_free += __popc(spare);
      __syncwarp();
      const int rank = __popc(need & lt_mask);
      if (!have && rank < n_resume) {
        my_slot = l_resume[n_resume - 1 - rank];
        load_ray(my_slot, true);
        have = true;
      }
      n_resume -= min(n_resume, __popc(need));
    }
    {
      // New rays.  The warp takes windows of kWindow consecutive rays (one camera, points that are
      // neighbours in Morton order), drops the ones whose point another camera already saw, and queues
      // the rest; lanes without a ray start the queued candidates.  All rays of a warp therefore stay
      // spatially close: their hash-grid and occupancy lookups share cache lines.
      const uint32_t need = __ballot_sync(0xffffffffu, !have);
      const int nn = min(__popc(need), n_free);
      for (int tries = 0; n_cand < nn && !global_done && tries < 4; ++tries) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(counter, (unsigned long long)kWindow);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= total) { global_done = true; break; }
        const int ci0 = (int)(base / (unsigned long long)n_act);
        const int j0 = (int)(base - (unsigned long long)ci0 * (unsigned long long)n_act);
#pragma unroll
        for (int k = 0; k < kWindow / 32; ++k) {
          bool ok = base + (unsigned long long)(k * 32 + lane) < total;
          int j = j0 + k * 32 + lane, ci = ci0;
          while (ok && j >= n_act) { j -= n_act; ++ci; }
          int pi = 0;
          if (ok) {
            pi = active_idx ? active_idx[j] : j;
            ok = surface[pi] == 0;                          // else another camera already saw this point
          }
          const uint32_t m = __ballot_sync(0xffffffffu, ok);
          if (ok) l_cand[n_cand + __popc(m & lt_mask)] = (uint32_t)pi | ((uint32_t)ci << kPiBits);
          n_cand += __popc(m);
        }
        __syncwarp();
      }
      const int take = min(nn, n_cand);
      const int rank = __popc(need & lt_mask);
      bool started = false;
      if (!have && rank < take) {
        const uint32_t pc = l_cand[n_cand - take + rank];
        const int pi = (int)(pc & ((1u << kPiBits) - 1u)), ci = (int)(pc >> kPiBits);
        float pt[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) { ray.o[d] = s_cams[ci * 3 + d]; pt[d] = points[pi * 3 + d]; }
        // unit direction, scene-box clip (nerfacc ray_aabb_intersect), t_max = |p - o|, first interval: march_math.h
        if (drb_ray_begin(ray.o, pt, a.scene_min, a.scene_max, a.step, ray.dir, ray.inv, &ray.len, &ray.t0, &ray.t1,
                          &ray.tm)) {
          ray.pi = pi; ray.ci = ci;
          ray.T = 1.f; ray.best = 0.f;
          started = true;
        }
      }
      n_cand -= take;
      // slots for the rays that really started (warp-uniform bookkeeping)
      const uint32_t st = __ballot_sync(0xffffffffu, started);
      if (started) { my_slot = l_free[n_free - 1 - __popc(st & lt_mask)]; have = true; }
      n_free -= __popc(st);
      __syncwarp();
      if (started) ++st_rays;
    }

    // ---------------- phase A: advance until a sample is pending (bounded number of skips) -------
    // occupancy of the cell that contains o + tm * dir; u = (x - roi_min) / extent (IEEE division, as
    // nerfacc's roi_to_unit) is handed back for the distance-to-next-voxel rule
    auto occupied_at = [&](float tm, float (&u)[3]) -> bool {
      float x[3];
      int idx[3];
      drb_sample_pos(tm, ray.dir, ray.o, x);
      if (!drb_voxel_of(x, a.roi_min, roi_ext, aux.roi_rcp, a.res, u, idx)) return false;
      const int cb = ((idx[0] >> aux.coarse_shift) * aux.coarse_dim + (idx[1] >> aux.coarse_shift)) *
                         aux.coarse_dim + (idx[2] >> aux.coarse_shift);
      if (!((s_coarse[cb >> 5] >> (cb & 31)) & 1u)) return false;
      return occ[((long long)idx[0] * a.res + idx[1]) * a.res + idx[2]] != 0;
    };
    bool pending = false;
    int nspec = 1;
    if (have) {
      int budget = kMaxSkips;
      while (budget > 0) {
        if (!(ray.tm < ray.len)) { have = false; break; }       // reached the point without a hit
        float u[3];
        if (occupied_at(ray.tm, u)) { pending = true; break; }
        --budget;
        drb_skip_empty(u, ray.dir, ray.inv, a.res, aux.res_rcp, roi_ext, a.step, &ray.t0, &ray.t1, &ray.tm);
        ++st_skips;
      }
      if (pending && kSpec > 1) {
        // Speculation along the ray: if the next sample positions (the chain t0 = t1, t1 = t0 + step,
        // tm = (t0 + t1) / 2 the marcher follows after a non-terminating sample) also lie in occupied
        // cells before the end of the ray, they are certain to be the next samples unless the ray
        // terminates first.  They are evaluated in the same batch: consecutive samples of one ray are
        // 1 step apart and share table sectors (the density phase is bound by distinct sectors per
        // load instruction), at the price of a few wasted samples when a ray terminates early.
        float t0s = ray.t0, t1s = ray.t1, tms = ray.tm;
        for (int j = 1; j < kSpec; ++j) {
          drb_chain_next(a.step, &t0s, &t1s, &tms);
          float u[3];
          if (!(tms < ray.len) || !occupied_at(tms, u)) break;
          ++nspec;
        }
      }
    }
    // ---------------- park the rays whose sample is pending --------------------------------------
    {
      const uint32_t pm = __ballot_sync(0xffffffffu, pending);
      if (pending) {
        store_ray(my_slot);
        l_pend[n_pend + __popc(pm & lt_mask)] = (uint16_t)(my_slot | ((nspec - 1) << 6));
        my_slot = -1;
        have = false;
      }
      n_pend += __popc(pm);
      __syncwarp();
    }
    // ---------------- phase B: kRaysPerBatch rays x up to kSpec consecutive samples at a time -----
    const uint32_t busy = __ballot_sync(0xffffffffu, have);
    const bool drain = global_done && busy == 0 && n_resume == 0 && n_cand == 0;
    if (n_pend == 0 && drain) break;
    while (n_pend >= kRaysPerBatch || (drain && n_pend > 0)) {
      // lanes that are in the middle of a ray park it in its home slot (registers are needed below)
      if (have) store_ray(my_slot);
      const int nb = min(kRaysPerBatch, n_pend);
      const int br = lane / kSpec, bj = lane % kSpec;          // ray of the batch, sample along the ray
      const uint32_t entry = br < nb ? (uint32_t)l_pend[n_pend - nb + br] : 0u;
      const int s = (int)(entry & 63u);
      const int ns = (int)(entry >> 6) + 1;                    // samples of this ray in the batch
      const bool valid = br < nb && bj < ns;
      n_pend -= nb;
      float xn[3] = {0.5f, 0.5f, 0.5f};
      bool inside = false;
      float dt = 0.f;
      if (valid) {
        const int ci = (int)(__float_as_uint(slots[9 * kSlots + s]) >> kPiBits);
        // sample bj of the chain: the stored (t0, t1, tm) for bj = 0, then t0 = t1, t1 = t0 + step
        float t0 = slots[4 * kSlots + s], t1 = slots[5 * kSlots + s], tm = slots[6 * kSlots + s];
        for (int jj = 0; jj < bj; ++jj) drb_chain_next(a.step, &t0, &t1, &tm);
        dt = t1 - t0;
        const float sdir[3] = {slots[0 * kSlots + s], slots[1 * kSlots + s], slots[2 * kSlots + s]};
        float x[3];
        drb_sample_pos(tm, sdir, &s_cams[ci * 3], x);
        float xt[3];
        inside = normalise(p, x, xt);
        if (inside) { xn[0] = xt[0]; xn[1] = xt[1]; xn[2] = xt[2]; }
        ++st_samples;
      }
      const float raw = warp_density_raw(p, s_lvl, mlp, xn, lane);
      const float sigma = (valid && inside) ? expf(raw - 1.f) : 0.f;
      const float alpha_mine = drb_alpha(sigma, dt);
      float alphas[kSpec];
#pragma unroll
      for (int jj = 0; jj < kSpec; ++jj) alphas[jj] = __shfl_sync(0xffffffffu, alpha_mine, (lane & ~(kSpec - 1)) + jj);
      bool resume = false, release = false;
      if (br < nb && bj == 0) {
        // the ray's samples in order, exactly as the sequential marcher would take them
        float t0 = slots[4 * kSlots + s], t1 = slots[5 * kSlots + s];
        float T = slots[7 * kSlots + s], best = slots[8 * kSlots + s];
        const int pi = (int)(__float_as_uint(slots[9 * kSlots + s]) & ((1u << kPiBits) - 1u));
        bool done = false;
#pragma unroll
        for (int jj = 0; jj < kSpec; ++jj) {
          if (jj < ns && !done) {
            const int verdict = drb_accumulate(alphas[jj], a.cut_off, &T, &best);    // march_math.h
            if (verdict == 1) surface[pi] = 1;
            if (verdict != 0) {
              done = true;
            } else if (T < a.cut_off || surface[pi]) {
              // exact early out: every later sample contributes alpha * T' <= T' <= T < cut_off
              done = true;
            }
            if (!done) { t0 = t1; t1 = t0 + a.step; }
          }
        }
        if (done) {
          release = true;
        } else {
          slots[4 * kSlots + s] = t0; slots[5 * kSlots + s] = t1; slots[6 * kSlots + s] = 0.5f * (t0 + t1);
          slots[7 * kSlots + s] = T; slots[8 * kSlots + s] = best;
          resume = true;
        }
      }
      const uint32_t rm = __ballot_sync(0xffffffffu, resume), fm = __ballot_sync(0xffffffffu, release);
      if (resume) l_resume[n_resume + __popc(rm & lt_mask)] = (uint8_t)s;
      if (release) l_free[n_free + __popc(fm & lt_mask)] = (uint8_t)s;
      n_resume += __popc(rm);
      n_free += __popc(fm);
      __syncwarp();
      load_ray(my_slot, have);
    }
  }
  {
    const unsigned long long r = __reduce_add_sync(0xffffffffu, st_rays), k = __reduce_add_sync(0xffffffffu, st_skips);
    const unsigned long long m = __reduce_add_sync(0xffffffffu, st_samples);
    if (lane == 0) {
      atomicAdd(&g_march_stats[0], r); atomicAdd(&g_march_stats[1], k);
      atomicAdd(&g_march_stats[2], m); atomicAdd(&g_march_stats[3], (unsigned long long)st_rounds);
    }
  }
}

// The stream-ordered allocator gives memory back to the OS at every synchronisation unless a release
// threshold is set; re-acquiring it costs milliseconds with a long tail.  Keep the pool.
static void keep_async_pool() {
  static bool done[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || done[dev]) return;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    unsigned long long thr = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
  }
  done[dev] = true;
}

struct MortonArgs { float roi_min[3], roi_inv[3]; };
__device__ __forceinline__ uint32_t spread10(uint32_t v) {      // 10 bits -> every third bit
  v = (v | (v << 16)) & 0x030000ffu;
  v = (v | (v << 8)) & 0x0300f00fu;
  v = (v | (v << 4)) & 0x030c30c3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}
// key = 30-bit Morton code of the point inside the ROI for active points, all ones otherwise (sorts last);
// *count = number of active points.  Only the ORDER of the rays depends on it, never a result.
__global__ void morton_key_kernel(const MortonArgs m, const float* __restrict__ points, int n,
                                  const uint8_t* __restrict__ active, uint32_t* __restrict__ keys,
                                  int* __restrict__ vals, int* __restrict__ count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  bool act = false;
  if (i < n) {
    act = active ? active[i] != 0 : true;
    uint32_t key = 0xffffffffu;
    if (act) {
      uint32_t q[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const float u = (points[i * 3 + d] - m.roi_min[d]) * m.roi_inv[d];
        q[d] = (uint32_t)fminf(fmaxf(u * 1024.f, 0.f), 1023.f);
      }
      key = spread10(q[0]) | (spread10(q[1]) << 1) | (spread10(q[2]) << 2);
    }
    keys[i] = key;
    vals[i] = i;
  }
  const uint32_t mask = __ballot_sync(0xffffffffu, act);
  if ((threadIdx.x & 31) == 0 && mask) atomicAdd(count, __popc(mask));
}

// Scratch of the extract entry points: carved from a caller workspace (the *_ws variants: no allocation inside the
// call) or taken from the stream-ordered allocator and given back when the call returns.
struct Arena {
  uint8_t* base = nullptr;
  size_t cap = 0, used = 0;
  cudaStream_t stream = nullptr;
  std::vector<void*> owned;
  Arena(void* ws, size_t bytes, cudaStream_t s) : base((uint8_t*)ws), cap(bytes), stream(s) {
    const uintptr_t mis = (uintptr_t)base & 255;
    if (base && mis) { const size_t skip = 256 - mis; base += skip; cap = cap > skip ? cap - skip : 0; }
  }
  void* take(size_t bytes) {
    bytes = (bytes + 255) & ~(size_t)255;
    if (base) {
      if (used + bytes > cap) return nullptr;
      void* p = base + used;
      used += bytes;
      return p;
    }
    void* p = nullptr;
    if (cudaMallocAsync(&p, bytes, stream) != cudaSuccess) return nullptr;
    owned.push_back(p);
    return p;
  }
  ~Arena() { for (void* p : owned) cudaFreeAsync(p, stream); }
};
static size_t march_scratch_bytes(int n, bool cellmajor) {
  if (n > (1 << 22) - 1) n = (1 << 22) - 1;            // larger point sets are marched in chunks
  size_t cub_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                  (const int*)nullptr, (int*)nullptr, n > 0 ? n : 1, 0, 32, nullptr);
  const size_t arr = ((size_t)(n > 0 ? n : 1) * 4 + 255) & ~(size_t)255;
  const LevelTable lv = host_levels();
  size_t cm_cells = 0;
  for (int l = 1; l <= 4; ++l) cm_cells += (size_t)lv.res[l] * lv.res[l] * lv.res[l];
  return 256 + 4 * arr + (size_t)kCoarseWords * 4 + cub_bytes + 256 + (cellmajor ? cm_cells * 64 : 0) + 1024;
}

// ev: optional {before, after, before2, after2} events recorded right around the marcher kernel itself (not the
// Morton sort / coarse bitmap / cell-major copy that precede it) - roofline instrumentation
static int surface_mask_chunk(const drb_ngp_params* pp, const uint8_t* occ_binary, int res,
                              const float* roi_aabb_host, const float* scene_aabb_host, const float* points,
                              int n, const float* cam_origins, int ncams, float step, float cut_off,
                              const uint8_t* active, uint8_t* surface, cudaStream_t stream,
                              cudaEvent_t* ev, Arena* arena_in);
// A ray word holds 22 bits of point index: larger point sets (a 256^3 block with more than a quarter of its cells
// occupied) are marched in chunks of points - the chunks are independent, the scratch is reused (same stream).
static int surface_mask_impl(const drb_ngp_params* pp, const uint8_t* occ_binary, int res,
                             const float* roi_aabb_host, const float* scene_aabb_host, const float* points,
                             int n, const float* cam_origins, int ncams, float step, float cut_off,
                             const uint8_t* active, uint8_t* surface, cudaStream_t stream,
                             cudaEvent_t* ev = nullptr, Arena* arena_in = nullptr) {
  const int kMaxPoints = (1 << 22) - 1;
  if (n <= kMaxPoints)
    return surface_mask_chunk(pp, occ_binary, res, roi_aabb_host, scene_aabb_host, points, n, cam_origins, ncams, step,
                              cut_off, active, surface, stream, ev, arena_in);
  for (int c0 = 0; c0 < n; c0 += kMaxPoints) {
    const int nn = n - c0 < kMaxPoints ? n - c0 : kMaxPoints;
    const size_t mark = arena_in ? arena_in->used : 0;
    const int rc = surface_mask_chunk(pp, occ_binary, res, roi_aabb_host, scene_aabb_host, points + 3 * (size_t)c0, nn,
                                      cam_origins, ncams, step, cut_off, active ? active + c0 : nullptr, surface + c0,
                                      stream, ev, arena_in);
    if (arena_in && arena_in->base) arena_in->used = mark;       // caller workspace: the next chunk reuses it
    if (rc) return rc;
  }
  return 0;
}
static int surface_mask_chunk(const drb_ngp_params* pp, const uint8_t* occ_binary, int res,
                              const float* roi_aabb_host, const float* scene_aabb_host, const float* points,
                              int n, const float* cam_origins, int ncams, float step, float cut_off,
                              const uint8_t* active, uint8_t* surface, cudaStream_t stream,
                              cudaEvent_t* ev, Arena* arena_in) {
  // empty inputs are legal (no point, or no camera: nothing is seen) and come with null data pointers
  DRB_REQUIRE(n >= 0 && ncams >= 0, "drb_surface_mask: negative count");
  DRB_REQUIRE(pp && occ_binary && roi_aabb_host && scene_aabb_host && (n == 0 || (points && surface)) &&
                  (ncams == 0 || cam_origins), "drb_surface_mask: null argument");
  DRB_REQUIRE(res > 0 && step > 0.f, "drb_surface_mask: bad grid / step");
  if (n > 0) DRB_CUDA_OK(cudaMemsetAsync(surface, 0, (size_t)n, stream));
  if (n == 0 || ncams == 0) {
    if (ev) for (int i = 0; i < 4; ++i) if (ev[i]) cudaEventRecord(ev[i], stream);     // nothing to march: zero-length bracket
    return 0;
  }
  const NgpDev p = make_dev(pp);
  MarchArgs a;
  for (int d = 0; d < 3; ++d) {
    a.roi_min[d] = roi_aabb_host[d]; a.roi_max[d] = roi_aabb_host[3 + d];
    a.scene_min[d] = scene_aabb_host[d]; a.scene_max[d] = scene_aabb_host[3 + d];
  }
  a.res = res; a.step = step; a.cut_off = cut_off;
  {
    static int skips = 0;
    if (!skips) { const char* env = getenv("DRB_MARCH_SKIPS"); skips = env ? atoi(env) : 16; if (skips < 1) skips = 16; }
    a.max_skips = skips;
    // DRB_MARCH_WATCHDOG_CLOCKS: test hook (a tiny limit forces the watchdog path)
    const char* wd = getenv("DRB_MARCH_WATCHDOG_CLOCKS");
    a.watchdog_clocks = wd ? atoll(wd) : 60000000000LL;
    a.watchdog_mask = wd ? 0x0u : 0xfffu;
    a.err = igemm_err_flag();
  }
  DRB_REQUIRE(levels_ok(p.lv), "drb_surface_mask: compile-time level table disagrees with host_levels()");
  DRB_REQUIRE(((uintptr_t)pp->hash_table & 15) == 0, "drb_surface_mask: hash table must be 16-byte aligned");
  DRB_REQUIRE(n < (1 << kPiBits) && ncams <= kMaxCams,
              "drb_surface_mask: at most %d points and %d cameras per call", (1 << kPiBits) - 1, kMaxCams);
  const size_t smem = march_smem_bytes(ncams);
  DRB_REQUIRE(smem <= 232448, "drb_surface_mask: %d cameras do not fit the shared-memory plan", ncams);
  DRB_CUDA_OK(cudaFuncSetAttribute(surface_mask_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  MarchAux aux;
  aux.coarse_shift = 0;
  while (((res + (1 << aux.coarse_shift) - 1) >> aux.coarse_shift) > kCoarseMaxDim) ++aux.coarse_shift;
  aux.coarse_dim = (res + (1 << aux.coarse_shift) - 1) >> aux.coarse_shift;
  aux.coarse_words = (aux.coarse_dim * aux.coarse_dim * aux.coarse_dim + 31) / 32;
  for (int d = 0; d < 3; ++d) {
    const float ext = a.roi_max[d] - a.roi_min[d];
    DRB_REQUIRE(ext > 0.f, "drb_surface_mask: empty ROI");
    // the reciprocal-multiply division is exact unless the divisor's significand is all ones
    uint32_t bits;
    memcpy(&bits, &ext, 4);
    DRB_REQUIRE((bits & 0x7fffffu) != 0x7fffffu, "drb_surface_mask: unsupported ROI extent");
    aux.roi_rcp[d] = (float)(1.0 / (double)ext);
  }
  aux.res_rcp = (float)(1.0 / (double)res);
  // scratch: ray counter, active count, the active points in Morton order (rays of neighbouring points
  // run together in one warp), coarse occupancy bitmap, CUB temporaries
  keep_async_pool();
  uint8_t* scratch = nullptr;
  size_t cub_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                  (const int*)nullptr, (int*)nullptr, n, 0, 32, stream);
  const size_t arr = ((size_t)n * 4 + 255) & ~(size_t)255;
  const size_t off_kin = 256, off_kout = off_kin + arr, off_vin = off_kout + arr, off_vout = off_vin + arr;
  const size_t off_coarse = off_vout + arr;
  const size_t off_cub = off_coarse + (size_t)kCoarseWords * 4;
  // cell-major records of the dense levels 1..4 (DRB_MARCH_CELLMAJOR=0 keeps the plain table layout)
  static int use_cm = -1;
  if (use_cm < 0) { const char* env = getenv("DRB_MARCH_CELLMAJOR"); use_cm = env ? atoi(env) : 1; }
  NgpDev pm = p;
  uint32_t cm_cells = 0;
  for (int l = 1; l <= 4; ++l) { pm.cm_off[l] = cm_cells; cm_cells += p.lv.res[l] * p.lv.res[l] * p.lv.res[l]; }
  const size_t off_cm = (off_cub + cub_bytes + 255) & ~(size_t)255;
  const size_t cm_bytes = use_cm ? (size_t)cm_cells * 64 : 0;
  Arena local(nullptr, 0, stream);
  Arena& arena = arena_in ? *arena_in : local;
  scratch = (uint8_t*)arena.take(off_cm + cm_bytes + 256);
  DRB_REQUIRE(scratch != nullptr, "drb_surface_mask: %s",
              arena.base ? "workspace too small (see the *_workspace_bytes query)" : "out of device memory");
  DRB_CUDA_OK(cudaMemsetAsync(scratch, 0, 256, stream));
  unsigned long long* counter = (unsigned long long*)scratch;
  int* count = (int*)(scratch + 64);
  int* idx = (int*)(scratch + off_vout);
  {
    MortonArgs ma;
    for (int d = 0; d < 3; ++d) { ma.roi_min[d] = a.roi_min[d]; ma.roi_inv[d] = 1.f / (a.roi_max[d] - a.roi_min[d]); }
    morton_key_kernel<<<cdiv(n, 256), 256, 0, stream>>>(ma, points, n, active, (uint32_t*)(scratch + off_kin),
                                                        (int*)(scratch + off_vin), count);
    DRB_LAUNCH_OK();
    DRB_CUDA_OK(cub::DeviceRadixSort::SortPairs(scratch + off_cub, cub_bytes, (const uint32_t*)(scratch + off_kin),
                                                (uint32_t*)(scratch + off_kout), (const int*)(scratch + off_vin),
                                                idx, n, 0, 32, stream));
  }
  aux.coarse = (const uint32_t*)(scratch + off_coarse);
  coarse_occ_kernel<<<cdiv(aux.coarse_words * 32, 256), 256, 0, stream>>>(occ_binary, res, aux.coarse_shift,
                                                                         aux.coarse_dim, (uint32_t*)(scratch + off_coarse));
  DRB_LAUNCH_OK();
  if (use_cm) {
    cellmajor_kernel<<<cdiv((long long)cm_cells * 8, 256), 256, 0, stream>>>(pm, (float2*)(scratch + off_cm));
    DRB_LAUNCH_OK();
    pm.cm = (const float4*)(scratch + off_cm);
  }
  const int grid = igemm_num_sms();
  const int* cnt = count;
  if (ev) { if (ev[0]) cudaEventRecord(ev[0], stream); if (ev[2]) cudaEventRecord(ev[2], stream); }
  surface_mask_kernel<<<grid, kMarchThreads, smem, stream>>>(pm, a, aux, occ_binary, points, n, cam_origins, ncams,
                                                             idx, cnt, counter, surface);
  if (ev) { if (ev[1]) cudaEventRecord(ev[1], stream); if (ev[3]) cudaEventRecord(ev[3], stream); }
  DRB_LAUNCH_OK();
  return 0;
}

extern "C" size_t drb_surface_mask_workspace_bytes(int n) { return march_scratch_bytes(n, true) + 512; }

extern "C" int drb_surface_mask_ws(const drb_ngp_params* pp, const uint8_t* occ_binary, int res,
                                   const float* roi_aabb_host, const float* scene_aabb_host, const float* points,
                                   int n, const float* cam_origins, int ncams, float step, float cut_off,
                                   uint8_t* surface, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  DRB_REQUIRE(workspace != nullptr, "drb_surface_mask_ws: null workspace");
  Arena arena(workspace, workspace_bytes, stream);
  return surface_mask_impl(pp, occ_binary, res, roi_aabb_host, scene_aabb_host, points, n, cam_origins, ncams,
                           step, cut_off, nullptr, surface, stream, nullptr, &arena);
}

extern "C" int drb_surface_mask(const drb_ngp_params* pp, const uint8_t* occ_binary, int res,
                                const float* roi_aabb_host, const float* scene_aabb_host, const float* points,
                                int n, const float* cam_origins, int ncams, float step, float cut_off,
                                uint8_t* surface, cudaStream_t stream) {
  return surface_mask_impl(pp, occ_binary, res, roi_aabb_host, scene_aabb_host, points, n, cam_origins, ncams,
                           step, cut_off, nullptr, surface, stream);
}

// ------------------------------------------------------------------------------------------
// A3 / A4 / A6: sample one jittered point per occupied cell, query, mask, scatter.
// ------------------------------------------------------------------------------------------
__global__ void sample_points_kernel(const long long* __restrict__ occupied, const float* __restrict__ jitter,
                                     int n, int res, const float* roi /* device copy not needed */,
                                     float rx0, float ry0, float rz0, float ex, float ey, float ez,
                                     float* __restrict__ points) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long idx = occupied[i];
  const int z = (int)(idx % res), y = (int)((idx / res) % res), x = (int)(idx / ((long long)res * res));
  // x = (coord + U[0,1)) / R (sample_grid.py:226-229), contract_inv AABB: x * (max - min) + min (:237-241)
  const float ux = __fdiv_rn((float)x + jitter[i * 3], (float)res);
  const float uy = __fdiv_rn((float)y + jitter[i * 3 + 1], (float)res);
  const float uz = __fdiv_rn((float)z + jitter[i * 3 + 2], (float)res);
  points[i * 3] = ux * ex + rx0;
  points[i * 3 + 1] = uy * ey + ry0;
  points[i * 3 + 2] = uz * ez + rz0;
  (void)roi;
}

__global__ void density_mask_kernel(const float* __restrict__ density, int n, float thre,
                                    float* __restrict__ alpha, uint8_t* __restrict__ dmask) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float d = density[i];
  // alpha = clip(1 - exp(-delta * density), 0, 1), delta = 1e-2 (sample_grid.py:112,341)
  alpha[i] = fminf(fmaxf(1.f - expf(-1e-2f * d), 0.f), 1.f);
  dmask[i] = d > thre ? 1 : 0;
}

__global__ void finish_extract_kernel(const long long* __restrict__ occupied, int n, const float* __restrict__ points,
                                      const float* __restrict__ rgb, const uint8_t* __restrict__ surface,
                                      const float* __restrict__ alpha, const uint8_t* __restrict__ dmask,
                                      float* __restrict__ grid) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float al = alpha[i];
  if (dmask[i] && surface[i] && grid) {
    float* g = grid + occupied[i] * 7;
    g[0] = points[i * 3]; g[1] = points[i * 3 + 1]; g[2] = points[i * 3 + 2];
    g[3] = rgb[i * 3]; g[4] = rgb[i * 3 + 1]; g[5] = rgb[i * 3 + 2];
    g[6] = al;
  }
}

// Roofline instrumentation: device time of the surface-field kernel (CUDA events on the launching
// stream).  drb_extract_last_surface_ms: the calling thread's most recent drb_extract_block.
// Profile mode: every drb_extract_block of this thread appends an event pair; drb_extract_read_profile
// synchronises on them once, returns the summed kernel time and the launch count, and clears the list
// (this is how bench.py times the kernel INSIDE its timed steps without a sync per call).
// idx[0 .. *count) = points with both masks set (any order); warp-aggregated atomics
__global__ void compact_masked_kernel(const uint8_t* __restrict__ dmask, const uint8_t* __restrict__ smask, int n,
                                      int* __restrict__ idx, int* __restrict__ count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool keep = i < n && dmask[i] != 0 && smask[i] != 0;
  const uint32_t m = __ballot_sync(0xffffffffu, keep);
  if (m == 0) return;
  const int lane = threadIdx.x & 31;
  int base = 0;
  if (lane == __ffs(m) - 1) base = atomicAdd(count, __popc(m));
  base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
  if (keep) idx[base + __popc(m & ((1u << lane) - 1u))] = i;
}

static thread_local cudaEvent_t g_surf_ev[2] = {nullptr, nullptr};
static thread_local bool g_surf_valid = false;
static thread_local bool g_surf_profile = false;
static thread_local std::vector<cudaEvent_t>* g_surf_list = nullptr;
extern "C" int drb_extract_last_surface_ms(float* host_ms) {
  DRB_REQUIRE(host_ms, "drb_extract_last_surface_ms: null argument");
  *host_ms = 0.f;
  if (!g_surf_valid) return 0;
  DRB_CUDA_OK(cudaEventSynchronize(g_surf_ev[1]));
  DRB_CUDA_OK(cudaEventElapsedTime(host_ms, g_surf_ev[0], g_surf_ev[1]));
  return 0;
}
extern "C" int drb_extract_set_profile(int on) {
  g_surf_profile = on != 0;
  if (g_surf_profile && !g_surf_list) g_surf_list = new std::vector<cudaEvent_t>();
  return 0;
}
extern "C" int drb_extract_read_profile(float* total_ms, int* launches) {
  DRB_REQUIRE(total_ms && launches, "drb_extract_read_profile: null argument");
  *total_ms = 0.f;
  *launches = 0;
  if (!g_surf_list) return 0;
  for (size_t i = 0; i + 1 < g_surf_list->size(); i += 2) {
    float ms = 0.f;
    DRB_CUDA_OK(cudaEventSynchronize((*g_surf_list)[i + 1]));
    DRB_CUDA_OK(cudaEventElapsedTime(&ms, (*g_surf_list)[i], (*g_surf_list)[i + 1]));
    *total_ms += ms;
    ++*launches;
  }
  for (cudaEvent_t e : *g_surf_list) cudaEventDestroy(e);
  g_surf_list->clear();
  return 0;
}

static int extract_block_impl(const drb_ngp_params* pp, const drb_extract_desc* e, float* points, float* rgb,
                              float* alpha, uint8_t* density_mask, uint8_t* surface_mask, float* voxel_grid,
                              Arena& arena, cudaStream_t stream);
extern "C" int drb_extract_block(const drb_ngp_params* pp, const drb_extract_desc* e, float* points, float* rgb,
                                 float* alpha, uint8_t* density_mask, uint8_t* surface_mask, float* voxel_grid,
                                 cudaStream_t stream) {
  keep_async_pool();
  Arena arena(nullptr, 0, stream);
  return extract_block_impl(pp, e, points, rgb, alpha, density_mask, surface_mask, voxel_grid, arena, stream);
}
extern "C" size_t drb_extract_workspace_bytes(int n_occupied) {
  const size_t n = (size_t)(n_occupied > 0 ? n_occupied : 1);
  auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
  return up(sizeof(float) * 15 * n) + up(sizeof(float) * n) + up(sizeof(int) * (n + 1)) +
         march_scratch_bytes(n_occupied, true) + 1024;
}
extern "C" int drb_extract_block_ws(const drb_ngp_params* pp, const drb_extract_desc* e, float* points, float* rgb,
                                    float* alpha, uint8_t* density_mask, uint8_t* surface_mask, float* voxel_grid,
                                    void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  DRB_REQUIRE(workspace != nullptr, "drb_extract_block_ws: null workspace");
  Arena arena(workspace, workspace_bytes, stream);
  return extract_block_impl(pp, e, points, rgb, alpha, density_mask, surface_mask, voxel_grid, arena, stream);
}
static int extract_block_impl(const drb_ngp_params* pp, const drb_extract_desc* e, float* points, float* rgb,
                              float* alpha, uint8_t* density_mask, uint8_t* surface_mask, float* voxel_grid,
                              Arena& arena, cudaStream_t stream) {
  DRB_REQUIRE(pp && e, "drb_extract_block: null argument");
  const int n = e->n_occupied;
  // a block without candidate cells is legal: zero grid, nothing else to write (the per-cell arrays may be null)
  DRB_REQUIRE(n >= 0 && (n == 0 || (points && rgb && alpha && density_mask && surface_mask)), "drb_extract_block: null argument");
  DRB_REQUIRE(n == 0 || (e->occupied && e->jitter && e->occ_binary && e->host_dirs && (e->ncams == 0 || e->cam_origins)),
              "drb_extract_block: null descriptor field");
  if (voxel_grid)
    DRB_CUDA_OK(cudaMemsetAsync(voxel_grid, 0, sizeof(float) * 7 * (size_t)e->res * e->res * e->res, stream));
  if (n == 0) return 0;
  sample_points_kernel<<<cdiv(n, 256), 256, 0, stream>>>(
      e->occupied, e->jitter, n, e->res, nullptr, e->roi_aabb[0], e->roi_aabb[1], e->roi_aabb[2],
      e->roi_aabb[3] - e->roi_aabb[0], e->roi_aabb[4] - e->roi_aabb[1], e->roi_aabb[5] - e->roi_aabb[2], points);
  DRB_LAUNCH_OK();
  // density / features go through scratch carved from the outputs: feat needs its own buffer
  float* feat = (float*)arena.take(sizeof(float) * 15 * (size_t)n);
  float* density = (float*)arena.take(sizeof(float) * (size_t)n);
  DRB_REQUIRE(feat && density, "drb_extract_block: %s",
              arena.base ? "workspace too small (drb_extract_workspace_bytes)" : "out of device memory");
  int rc = drb_ngp_density(pp, points, n, density, feat, stream);
  if (!rc) {
    density_mask_kernel<<<cdiv(n, 256), 256, 0, stream>>>(density, n, e->density_thre, alpha, density_mask);
    if (cudaGetLastError() != cudaSuccess) rc = DRB_ECUDA;
  }
  const bool rgb_late = e->rgb_only_where_masked != 0;     // colour only for the cells both masks keep
  if (!rc && !rgb_late) rc = drb_ngp_rgb_mean(pp, feat, n, e->host_dirs, e->ndirs, rgb, stream);
  if (!rc) {
    if (!g_surf_ev[0]) { cudaEventCreate(&g_surf_ev[0]); cudaEventCreate(&g_surf_ev[1]); }
    cudaEvent_t pe[2] = {nullptr, nullptr};
    if (g_surf_profile) { cudaEventCreate(&pe[0]); cudaEventCreate(&pe[1]); }
    cudaEvent_t evs[4] = {g_surf_ev[0], g_surf_ev[1], pe[0], pe[1]};
    rc = surface_mask_impl(pp, e->occ_binary, e->res, e->roi_aabb, e->scene_aabb, points, n, e->cam_origins,
                           e->ncams, e->render_step_size, e->cut_off,
                           e->surface_only_where_dense ? density_mask : nullptr, surface_mask, stream, evs, &arena);
    if (g_surf_profile) { g_surf_list->push_back(pe[0]); g_surf_list->push_back(pe[1]); }
    g_surf_valid = true;
  }
  if (!rc && rgb_late) {
    int* idx = (int*)arena.take(sizeof(int) * ((size_t)n + 1));
    if (!idx) rc = DRB_ENOMEM;
    if (!rc) {
      int* count = idx + n;
      cudaMemsetAsync(count, 0, sizeof(int), stream);
      cudaMemsetAsync(rgb, 0, sizeof(float) * 3 * (size_t)n, stream);
      compact_masked_kernel<<<cdiv(n, 256), 256, 0, stream>>>(density_mask, surface_mask, n, idx, count);
      if (cudaGetLastError() != cudaSuccess) rc = DRB_ECUDA;
      if (!rc) rc = rgb_mean_impl(pp, feat, n, e->host_dirs, e->ndirs, rgb, idx, count, stream);
    }
  }
  if (!rc) {
    finish_extract_kernel<<<cdiv(n, 256), 256, 0, stream>>>(e->occupied, n, points, rgb, surface_mask, alpha,
                                                           density_mask, voxel_grid);
    if (cudaGetLastError() != cudaSuccess) rc = DRB_ECUDA;
  }
  return rc;       // stream-ordered scratch goes back with the arena
}

}  // namespace drb
