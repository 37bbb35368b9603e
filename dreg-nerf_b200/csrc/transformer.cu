// R6 / R7 glue around the tensor-core GEMMs: LayerNorm(256) (+ positional add, plane split), the
// multi-head attention core (fp32, flash style: no N x S matrix is materialised, unlike the
// reference's bmm -> softmax -> bmm path of nn.MultiheadAttention), the decoder's
// softmax-weighted coordinate sum and the overlap head.
#include "common.cuh"

namespace drb {

// ------------------------------------------------------------------------------------------
// LayerNorm over 256 channels, one warp per row (8 values per lane), eps 1e-5.
// ------------------------------------------------------------------------------------------
__global__ void layernorm256_kernel(const float* __restrict__ x, int n,
                                    const float* __restrict__ gamma, const float* __restrict__ beta,
                                    const float* __restrict__ add, float* __restrict__ out,
                                    plane_t* __restrict__ out_hi, plane_t* __restrict__ out_lo) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const float* xr = x + (long long)row * 256;
  const int c0 = lane * 4, c1 = 128 + lane * 4;
  const float4 a = *(const float4*)(xr + c0);
  const float4 b = *(const float4*)(xr + c1);
  float s = a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w;
  s = warp_sum(s);
  const float mean = s * (1.f / 256.f);
  float v[8] = {a.x - mean, a.y - mean, a.z - mean, a.w - mean,
                b.x - mean, b.y - mean, b.z - mean, b.w - mean};
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) ss += v[i] * v[i];
  ss = warp_sum(ss);
  const float rstd = rsqrtf(ss * (1.f / 256.f) + 1e-5f);
  const float4 g0 = *(const float4*)(gamma + c0), g1 = *(const float4*)(gamma + c1);
  const float4 b0 = *(const float4*)(beta + c0), b1 = *(const float4*)(beta + c1);
  const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
  const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = v[i] * rstd * gg[i] + bb[i];
  if (add) {
    const float4 p0 = *(const float4*)(add + (long long)row * 256 + c0);
    const float4 p1 = *(const float4*)(add + (long long)row * 256 + c1);
    v[0] += p0.x; v[1] += p0.y; v[2] += p0.z; v[3] += p0.w;
    v[4] += p1.x; v[5] += p1.y; v[6] += p1.z; v[7] += p1.w;
  }
  if (out) {
    *(float4*)(out + (long long)row * 256 + c0) = make_float4(v[0], v[1], v[2], v[3]);
    *(float4*)(out + (long long)row * 256 + c1) = make_float4(v[4], v[5], v[6], v[7]);
  }
  if (out_hi) {
    const bool pair = out_lo != nullptr;
    plane_t h[8], l[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) split16(v[i], pair, h[i], l[i]);
    *(uint2*)(out_hi + (long long)row * 256 + c0) = make_uint2(pack16x2(h[0], h[1]), pack16x2(h[2], h[3]));
    *(uint2*)(out_hi + (long long)row * 256 + c1) = make_uint2(pack16x2(h[4], h[5]), pack16x2(h[6], h[7]));
    if (out_lo) {
      *(uint2*)(out_lo + (long long)row * 256 + c0) = make_uint2(pack16x2(l[0], l[1]), pack16x2(l[2], l[3]));
      *(uint2*)(out_lo + (long long)row * 256 + c1) = make_uint2(pack16x2(l[4], l[5]), pack16x2(l[6], l[7]));
    }
  }
}

extern "C" int drb_layernorm256(const float* x, int n, const float* gamma, const float* beta,
                                const float* add, float* out, void* out_hi, void* out_lo,
                                cudaStream_t stream) {
  DRB_REQUIRE(x && gamma && beta && (out || out_hi), "drb_layernorm256: bad arguments");
  if (n == 0) return 0;
  layernorm256_kernel<<<cdiv(n, 8), 256, 0, stream>>>(x, n, gamma, beta, add, out, (plane_t*)out_hi,
                                                      (plane_t*)out_lo);
  DRB_LAUNCH_OK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// Attention core, head dim 32: flash-attention style on the tensor cores.  Block = 4 warps = one
// 16-query tile of one head; K / V go through shared memory in 64-key tiles (row pitch 36 words:
// conflict-free fragment reads), every warp owns a quarter (16 keys) of each tile with its own online
// soft-max state, and the four partial (max, sum, O) are merged through shared memory at the end.
// Both products run as 3xTF32 mma.sync m16n8k8 (hi*hi + lo*hi + hi*lo, fp32 accumulate, ~2^-21
// relative - the fp32-grade contract of the register path): S = (Q * scale) K^T, then P = exp2(S - m)
// in the accumulator layout is fed straight back as the A operand of P V by permuting the key index
// inside each block of 8 (fragment column t <-> key 2t, t+4 <-> key 2t+1; the V fragments are read
// with the same permutation), so P never leaves registers.
// ------------------------------------------------------------------------------------------
static constexpr int kHd = 32;
static constexpr int kKeyTile = 64;
static constexpr int kKvPitch = 36;

__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  const float h = tf32_hi(x);
  hi = __float_as_uint(h);
  lo = __float_as_uint(tf32_hi(x - h));
}
__device__ __forceinline__ void mma_1688(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma3(float (&c)[4], const uint32_t (&ahi)[4], const uint32_t (&alo)[4],
                                     float b0, float b1) {
  uint32_t h0, l0, h1, l1;
  split_tf32(b0, h0, l0);
  split_tf32(b1, h1, l1);
  mma_1688(c, alo, h0, h1);
  mma_1688(c, ahi, l0, l1);
  mma_1688(c, ahi, h0, h1);
}

__global__ void __launch_bounds__(128)
mha_core_kernel(const float* __restrict__ q, int ldq, const float* __restrict__ k, int ldk,
                const float* __restrict__ v, int ldv, int nq, int nk, float scale_log2,
                float* __restrict__ out, plane_t* __restrict__ out_hi, plane_t* __restrict__ out_lo,
                int ld_out) {
  __shared__ __align__(16) float sk[kKeyTile][kKvPitch];
  __shared__ __align__(16) float sv[kKeyTile][kKvPitch];
  __shared__ float sm[4][16], sl[4][16];
  __shared__ float so[4][16][kHd + 1];

  const int head = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int q0 = blockIdx.x * 16;

  // Q fragments (scaled): rows g / g+8 of the tile, columns 8*ks + t / + t + 4
  uint32_t qhi[4][4], qlo[4][4];
#pragma unroll
  for (int ks = 0; ks < 4; ++ks)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = q0 + g + 8 * (i & 1), c = 8 * ks + t + 4 * (i >> 1);
      const float x = r < nq ? q[(long long)r * ldq + head * kHd + c] * scale_log2 : 0.f;
      split_tf32(x, qhi[ks][i], qlo[ks][i]);
    }
  float o[4][4];                         // O tile: 4 n-tiles of 8 dims, accumulator layout
#pragma unroll
  for (int nt = 0; nt < 4; ++nt)
#pragma unroll
    for (int i = 0; i < 4; ++i) o[nt][i] = 0.f;
  float m[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f};     // rows g, g + 8 (l: this lane's share)

  for (int k0 = 0; k0 < nk; k0 += kKeyTile) {
    __syncthreads();
    for (int i = threadIdx.x; i < kKeyTile * (kHd / 4); i += 128) {
      const int r = i / (kHd / 4), c = (i % (kHd / 4)) * 4;
      float4 kk = make_float4(0.f, 0.f, 0.f, 0.f), vv = kk;
      if (k0 + r < nk) {
        kk = *(const float4*)(k + (long long)(k0 + r) * ldk + head * kHd + c);
        vv = *(const float4*)(v + (long long)(k0 + r) * ldv + head * kHd + c);
      }
      *(float4*)&sk[r][c] = kk;
      *(float4*)&sv[r][c] = vv;
    }
    __syncthreads();
    const int kb = warp * 16;                         // this warp's 16 keys of the tile
    if (k0 + kb >= nk) continue;                      // (warp-uniform) nothing valid in this quarter
    // ---- S = Q K^T for 2 n-tiles of 8 keys ----
    float sacc[2][4];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
#pragma unroll
      for (int i = 0; i < 4; ++i) sacc[j][i] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const float* kr = &sk[kb + 8 * j + g][8 * ks + t];    // B fragment: (k = dim, n = key g)
        mma3(sacc[j], qhi[ks], qlo[ks], kr[0], kr[4]);
      }
      // accumulator (row g | g+8, key 2t | 2t+1): mask keys beyond nk
      const int key = k0 + kb + 8 * j + 2 * t;
      if (key >= nk) { sacc[j][0] = -INFINITY; sacc[j][2] = -INFINITY; }
      if (key + 1 >= nk) { sacc[j][1] = -INFINITY; sacc[j][3] = -INFINITY; }
    }
    // ---- online soft-max over the 16 keys, rows g (regs 0, 1) and g + 8 (regs 2, 3) ----
    float corr[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      float mx = fmaxf(fmaxf(sacc[0][2 * r], sacc[0][2 * r + 1]), fmaxf(sacc[1][2 * r], sacc[1][2 * r + 1]));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      mx = fmaxf(mx, m[r]);                           // finite: the first key of the quarter is valid
      corr[r] = exp2f(m[r] - mx);                     // m == -inf on the first group -> 0
      m[r] = mx;
      l[r] *= corr[r];
    }
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      o[nt][0] *= corr[0]; o[nt][1] *= corr[0];
      o[nt][2] *= corr[1]; o[nt][3] *= corr[1];
    }
    // ---- O += P V, P straight from the accumulator (keys permuted inside each block of 8) ----
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const float p0 = exp2f(sacc[j][0] - m[0]), p1 = exp2f(sacc[j][1] - m[0]);
      const float p2 = exp2f(sacc[j][2] - m[1]), p3 = exp2f(sacc[j][3] - m[1]);
      l[0] += p0 + p1;
      l[1] += p2 + p3;
      uint32_t phi[4], plo[4];
      split_tf32(p0, phi[0], plo[0]);                 // a0 = (row g,   key 2t)
      split_tf32(p2, phi[1], plo[1]);                 // a1 = (row g+8, key 2t)
      split_tf32(p1, phi[2], plo[2]);                 // a2 = (row g,   key 2t+1)
      split_tf32(p3, phi[3], plo[3]);                 // a3 = (row g+8, key 2t+1)
      const float* v0 = &sv[kb + 8 * j + 2 * t][g];   // B fragment: (k = key 2t | 2t+1, n = dim 8*nt + g)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) mma3(o[nt], phi, plo, v0[8 * nt], v0[kKvPitch + 8 * nt]);
    }
  }
  // ---- merge the 4 key splits ----
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l[r] += __shfl_xor_sync(0xffffffffu, l[r], 1);
    l[r] += __shfl_xor_sync(0xffffffffu, l[r], 2);
  }
  if (t == 0) {
    sm[warp][g] = m[0]; sm[warp][g + 8] = m[1];
    sl[warp][g] = l[0]; sl[warp][g + 8] = l[1];
  }
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    so[warp][g][8 * nt + 2 * t] = o[nt][0]; so[warp][g][8 * nt + 2 * t + 1] = o[nt][1];
    so[warp][g + 8][8 * nt + 2 * t] = o[nt][2]; so[warp][g + 8][8 * nt + 2 * t + 1] = o[nt][3];
  }
  __syncthreads();
  // thread: query row = tid / 8, dims (tid % 8) * 4 .. + 4
  const int mq = threadIdx.x >> 3, md = (threadIdx.x & 7) * 4;
  const int oq = q0 + mq;
  if (oq < nq) {
    const float mm = fmaxf(fmaxf(sm[0][mq], sm[1][mq]), fmaxf(sm[2][mq], sm[3][mq]));
    float w[4], lt = 0.f;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      w[u] = (sm[u][mq] == -INFINITY) ? 0.f : exp2f(sm[u][mq] - mm);
      lt += w[u] * sl[u][mq];
    }
    const float inv = 1.f / lt;
    float r[4];
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      float acc = 0.f;
#pragma unroll
      for (int u = 0; u < 4; ++u) acc += w[u] * so[u][mq][md + d];
      r[d] = acc * inv;
    }
    const long long off = (long long)oq * ld_out + head * kHd + md;
    if (out) *(float4*)(out + off) = make_float4(r[0], r[1], r[2], r[3]);
    if (out_hi) {
      const bool pair = out_lo != nullptr;
      plane_t h[4], lo4[4];
#pragma unroll
      for (int d = 0; d < 4; ++d) split16(r[d], pair, h[d], lo4[d]);
      *(uint2*)(out_hi + off) = make_uint2(pack16x2(h[0], h[1]), pack16x2(h[2], h[3]));
      if (out_lo) *(uint2*)(out_lo + off) = make_uint2(pack16x2(lo4[0], lo4[1]), pack16x2(lo4[2], lo4[3]));
    }
  }
}

extern "C" int drb_mha_core(const float* q, int ldq, const float* k, int ldk, const float* v,
                            int ldv, int nq, int nk, int heads, float scale, float* out,
                            void* out_hi, void* out_lo, int ld_out, cudaStream_t stream) {
  DRB_REQUIRE(q && k && v && (out || out_hi), "drb_mha_core: null argument");
  DRB_REQUIRE(ldq % 4 == 0 && ldk % 4 == 0 && ldv % 4 == 0 && ld_out % 8 == 0,
              "drb_mha_core: pitches must be multiples of 4 (inputs) / 8 (output)");
  DRB_REQUIRE(nk > 0, "drb_mha_core: empty key set");
  if (nq == 0) return 0;
  const float scale_log2 = scale * 1.4426950408889634f;
  dim3 grid((unsigned)cdiv(nq, 16), (unsigned)heads);
  mha_core_kernel<<<grid, 128, 0, stream>>>(q, ldq, k, ldk, v, ldv, nq, nk, scale_log2, out,
                                            (plane_t*)out_hi, (plane_t*)out_lo, ld_out);
  DRB_LAUNCH_OK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// Decoder tail: out[i] = sum_j softmax_j(s[i][j]) * xyz[j].  One warp per query row.
// ------------------------------------------------------------------------------------------
__global__ void softmax_xyz_kernel(const float* __restrict__ s, int ld, int nq, int nk,
                                   const float* __restrict__ xyz, int ld_xyz,
                                   float* __restrict__ out) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= nq) return;
  const float* sr = s + (long long)row * ld;
  float mx = -INFINITY;
  for (int j = lane; j < nk; j += 32) mx = fmaxf(mx, sr[j]);
  mx = warp_max(mx);
  float l = 0.f, ax = 0.f, ay = 0.f, az = 0.f;
  for (int j = lane; j < nk; j += 32) {
    const float p = expf(sr[j] - mx);
    l += p;
    const float* pj = xyz + (long long)j * ld_xyz;
    ax = fmaf(p, pj[0], ax); ay = fmaf(p, pj[1], ay); az = fmaf(p, pj[2], az);
  }
  l = warp_sum(l); ax = warp_sum(ax); ay = warp_sum(ay); az = warp_sum(az);
  if (lane == 0) {
    out[(long long)row * 3 + 0] = ax / l;
    out[(long long)row * 3 + 1] = ay / l;
    out[(long long)row * 3 + 2] = az / l;
  }
}

extern "C" int drb_softmax_weighted_xyz(const float* s, int ld, int nq, int nk, const float* xyz,
                                        int ld_xyz, float* out, cudaStream_t stream) {
  DRB_REQUIRE(s && xyz && out && nk > 0 && ld >= nk && ld_xyz >= 3,
              "drb_softmax_weighted_xyz: bad arguments");
  if (nq == 0) return 0;
  softmax_xyz_kernel<<<cdiv(nq, 8), 256, 0, stream>>>(s, ld, nq, nk, xyz, ld_xyz, out);
  DRB_LAUNCH_OK();
  return 0;
}

__global__ void overlap_kernel(const float* __restrict__ feat, int n, const float* __restrict__ w,
                               const float* __restrict__ b, float* __restrict__ out) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const float* f = feat + (long long)row * 256;
  float acc = 0.f;
#pragma unroll
  for (int c = lane * 4; c < 256; c += 128) {
    const float4 a = *(const float4*)(f + c);
    const float4 ww = *(const float4*)(w + c);
    acc += a.x * ww.x + a.y * ww.y + a.z * ww.z + a.w * ww.w;
  }
  acc = warp_sum(acc);
  if (lane == 0) out[row] = 1.f / (1.f + expf(-(acc + b[0])));
}

extern "C" int drb_overlap_sigmoid(const float* feat, int n, const float* w, const float* b,
                                   float* out, cudaStream_t stream) {
  DRB_REQUIRE(feat && w && b && out, "drb_overlap_sigmoid: null argument");
  if (n == 0) return 0;
  overlap_kernel<<<cdiv(n, 8), 256, 0, stream>>>(feat, n, w, b, out);
  DRB_LAUNCH_OK();
  return 0;
}

}  // namespace drb
