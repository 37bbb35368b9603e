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
// Attention core, head dim 32, fp32.  Block = 4 warps = 32 queries of one head; every lane owns one
// query (q and o in registers), the 4 warps split each 64-key shared-memory tile, partial
// (max, sum, o) are merged through shared memory at the end.  All lanes of a warp read the same
// K/V row (broadcast float4 loads).
// ------------------------------------------------------------------------------------------
static constexpr int kHd = 32;
static constexpr int kKeyTile = 64;

__global__ void __launch_bounds__(128)
mha_core_kernel(const float* __restrict__ q, int ldq, const float* __restrict__ k, int ldk,
                const float* __restrict__ v, int ldv, int nq, int nk, float scale_log2,
                float* __restrict__ out, plane_t* __restrict__ out_hi, plane_t* __restrict__ out_lo,
                int ld_out) {
  __shared__ __align__(16) float sk[kKeyTile][kHd];
  __shared__ __align__(16) float sv[kKeyTile][kHd];
  __shared__ float sm[4][32], sl[4][32];
  __shared__ float so[4][32][kHd + 1];

  const int head = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qi = blockIdx.x * 32 + lane;
  const bool q_ok = qi < nq;

  float qr[kHd], o[kHd];
  {
    const float* qp = q + (long long)(q_ok ? qi : 0) * ldq + head * kHd;
#pragma unroll
    for (int d = 0; d < kHd; d += 4) {
      const float4 t = *(const float4*)(qp + d);
      qr[d] = t.x * scale_log2; qr[d + 1] = t.y * scale_log2;
      qr[d + 2] = t.z * scale_log2; qr[d + 3] = t.w * scale_log2;
    }
#pragma unroll
    for (int d = 0; d < kHd; ++d) o[d] = 0.f;
  }
  float m = -INFINITY, l = 0.f;

  for (int k0 = 0; k0 < nk; k0 += kKeyTile) {
    __syncthreads();
    // cooperative tile load: 64 rows x 32 floats for K and V
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
    const int jend = min(kKeyTile, nk - k0);
    // this warp's quarter of the tile, 4 keys at a time
    for (int j0 = warp * 16; j0 < min(jend, warp * 16 + 16); j0 += 4) {
      float s[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float acc = 0.f;
#pragma unroll
        for (int d = 0; d < kHd; d += 4) {
          const float4 kk = *(const float4*)&sk[j0 + u][d];
          acc = fmaf(qr[d], kk.x, acc); acc = fmaf(qr[d + 1], kk.y, acc);
          acc = fmaf(qr[d + 2], kk.z, acc); acc = fmaf(qr[d + 3], kk.w, acc);
        }
        s[u] = (j0 + u < jend) ? acc : -INFINITY;
      }
      const float mx = fmaxf(fmaxf(fmaxf(s[0], s[1]), fmaxf(s[2], s[3])), m);
      const float corr = exp2f(m - mx);   // m == -inf on the first group -> 0
      l *= corr;
#pragma unroll
      for (int d = 0; d < kHd; ++d) o[d] *= corr;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float p = exp2f(s[u] - mx);
        l += p;
#pragma unroll
        for (int d = 0; d < kHd; d += 4) {
          const float4 vv = *(const float4*)&sv[j0 + u][d];
          o[d] = fmaf(p, vv.x, o[d]); o[d + 1] = fmaf(p, vv.y, o[d + 1]);
          o[d + 2] = fmaf(p, vv.z, o[d + 2]); o[d + 3] = fmaf(p, vv.w, o[d + 3]);
        }
      }
      m = mx;
    }
  }
  // merge the 4 key splits
  sm[warp][lane] = m;
  sl[warp][lane] = l;
#pragma unroll
  for (int d = 0; d < kHd; ++d) so[warp][lane][d] = o[d];
  __syncthreads();
  // thread t: query = t / 4, dims (t % 4) * 8 .. + 8
  const int mq = threadIdx.x >> 2, md = (threadIdx.x & 3) * 8;
  const int oq = blockIdx.x * 32 + mq;
  if (oq < nq) {
    float mm = fmaxf(fmaxf(sm[0][mq], sm[1][mq]), fmaxf(sm[2][mq], sm[3][mq]));
    float w[4], lt = 0.f;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      w[u] = (sm[u][mq] == -INFINITY) ? 0.f : exp2f(sm[u][mq] - mm);
      lt += w[u] * sl[u][mq];
    }
    const float inv = 1.f / lt;
    float r[8];
#pragma unroll
    for (int d = 0; d < 8; ++d) {
      float acc = 0.f;
#pragma unroll
      for (int u = 0; u < 4; ++u) acc += w[u] * so[u][mq][md + d];
      r[d] = acc * inv;
    }
    const long long off = (long long)oq * ld_out + head * kHd + md;
    if (out) {
      *(float4*)(out + off) = make_float4(r[0], r[1], r[2], r[3]);
      *(float4*)(out + off + 4) = make_float4(r[4], r[5], r[6], r[7]);
    }
    if (out_hi) {
      const bool pair = out_lo != nullptr;
      plane_t h[8], lo8[8];
#pragma unroll
      for (int d = 0; d < 8; ++d) split16(r[d], pair, h[d], lo8[d]);
      *(uint4*)(out_hi + off) = make_uint4(pack16x2(h[0], h[1]), pack16x2(h[2], h[3]),
                                           pack16x2(h[4], h[5]), pack16x2(h[6], h[7]));
      if (out_lo)
        *(uint4*)(out_lo + off) = make_uint4(pack16x2(lo8[0], lo8[1]), pack16x2(lo8[2], lo8[3]),
                                             pack16x2(lo8[4], lo8[5]), pack16x2(lo8[6], lo8[7]));
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
  dim3 grid((unsigned)cdiv(nq, 32), (unsigned)heads);
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
