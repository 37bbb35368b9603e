// Multi-head attention core (head dim 32) as a FlashAttention-style tcgen05 kernel: replaces the bmm -> softmax ->
// bmm of nn.MultiheadAttention inside TransformerCrossEncoderLayer.forward_pre
// (conerf/register/transformer.py:225-299) for self- and cross-attention.
//
//   pack   : q / k / v fp32 rows [n][ld] -> 16-bit planes per head: Q' = q * scale * log2(e) and K as
//            [head][token][64] (head dim zero-padded to one 128-byte swizzle row), V transposed as
//            [head][32][token] (keys contiguous: K-major B operand of P V)
//   kernel : one CTA per (head, 128-query tile).  warp 0 = TMA producer (Q once, K / V^T tiles of 128 keys, two
//            stages), warp 1 = tcgen05.mma issuer, warps 2-5 = soft-max: each thread owns one query row,
//            reads its 128 logits from TMEM (tcgen05.ld), keeps the online max / sum, writes P as a swizzled
//            K-major A operand into shared memory, and after P V^T (accumulated per tile in TMEM) folds the
//            tile's 32 output columns into its registers with the running rescale.
//            S = Q' K^T: 128 x 128 x 64;  O_tile = P V: 128 x 32 x 128.  S and P never touch HBM.
// Precision follows the GEMM planes: bf16, or fp16 hi/lo pairs with three MMAs per product (fp32-grade).
#include "common.cuh"

#include <stdlib.h>

namespace drb {

static constexpr int kAttThreads = 320;      // warp 0 TMA, warp 1 MMA, warps 2-9 soft-max (two threads per query row)
static constexpr int kAttQ = 128;            // queries per CTA == TMEM lanes
static constexpr int kAttK = 128;            // keys per tile
static constexpr uint32_t kQBytes = kAttQ * 64 * 2;          // 16 KB per plane
static constexpr uint32_t kKBytes = kAttK * 64 * 2;          // 16 KB per plane
static constexpr uint32_t kVBytes = 32 * kAttK * 2;          //  8 KB per plane (two [32][64] sub-tiles)
static constexpr uint32_t kPBytes = kAttQ * kAttK * 2;       // 32 KB per plane (two [128][64] sub-tiles)

struct AttProblem {
  int nq, nk;               // valid queries / keys
  int q_row0, k_row0;       // first row of the query / key segment inside the packed planes
  int out_row0;             // output row of query 0
};
struct AttArgs {
  AttProblem pr[2];         // blockIdx.z selects the problem (both directions of one attention in one launch)
  int heads, planes;
  float* out;               // fp32 [.][ld_out] or null
  plane_t* out_hi;
  plane_t* out_lo;
  int ld_out;
  int* err;
};

__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
// 2^x for x <= 0 on the MUFU (rel. error 2^-22; below -126 the result flushes to 0, which is what a soft-max wants)
__device__ __forceinline__ float ex2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// p = 2^(s - mx) of one thread's 64 logits -> 16-bit planes in shared memory (8 swizzled 16-byte chunks); returns the sum.
template <int P, bool kFull>
__device__ __forceinline__ float softmax_half_row(const float (&sv_)[64], float mx, int valid, uint32_t prow, int row) {
  float ls[4] = {0.f, 0.f, 0.f, 0.f};                 // four partial sums: no 32-deep add chain
#pragma unroll
  for (int c8 = 0; c8 < 8; ++c8) {
    uint32_t hw[4], lw[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j0 = c8 * 8 + 2 * u;
      float p0 = ex2_fast(sv_[j0] - mx), p1 = ex2_fast(sv_[j0 + 1] - mx);
      if (!kFull) { p0 = j0 < valid ? p0 : 0.f; p1 = j0 + 1 < valid ? p1 : 0.f; }
      ls[u] += p0 + p1;
      if (P == 1) {
        const __nv_bfloat162 h2 = __floats2bfloat162_rn(p0, p1);
        hw[u] = *(const uint32_t*)&h2;
        lw[u] = 0u;
      } else {
        const __half2 h2 = __floats2half2_rn(p0, p1);
        const float2 hf = __half22float2(h2);
        const __half2 l2 = __floats2half2_rn(p0 - hf.x, p1 - hf.y);
        hw[u] = *(const uint32_t*)&h2;
        lw[u] = *(const uint32_t*)&l2;
      }
    }
    const uint32_t dst = prow + (uint32_t)((c8 ^ (row & 7)) * 16);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(hw[0]), "r"(hw[1]), "r"(hw[2]), "r"(hw[3]) : "memory");
    if (P == 2)
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + kPBytes), "r"(lw[0]), "r"(lw[1]), "r"(lw[2]), "r"(lw[3]) : "memory");
  }
  return (ls[0] + ls[1]) + (ls[2] + ls[3]);
}

// Schedule of one CTA (head, 128 queries), tiles of 128 keys:
//   MMA warp : S(0); for t: S(t+1) as soon as every soft-max thread has read S(t); P V(t) when P(t) is in smem.
//   soft-max : 8 warps, the two threads of a row (warps w and w + 4 share a TMEM lane quarter) take 64 keys each:
//              read S from TMEM, exchange the row maximum through shared memory (named barrier of the warp pair),
//              exp2 on the MUFU, write their half of P (swizzled K-major A operand), and fold the P V product of
//              the PREVIOUS tile into their 16 output columns - P and the P V accumulator are double buffered
//              (bf16; the fp16-pair planes only fit one P buffer), so the soft-max never waits for the tensor core.
//   Roofline of this kernel is the MUFU (128 x 128 exp2 per tile = 1024 clk per SM at 16 / clk) - the two MMAs of a
//   tile take ~390 clk - so the tensor pipe cannot exceed ~35 % here whatever the schedule (head dim 32).
template <int P>
__global__ void __launch_bounds__(kAttThreads, 1)
att_fwd_kernel(const __grid_constant__ CUtensorMap tmQ0, const __grid_constant__ CUtensorMap tmQ1,
               const __grid_constant__ CUtensorMap tmK0, const __grid_constant__ CUtensorMap tmK1,
               const __grid_constant__ CUtensorMap tmV0, const __grid_constant__ CUtensorMap tmV1, const AttArgs args) {
  const AttProblem a = args.pr[blockIdx.z];
  if ((int)blockIdx.x * kAttQ >= a.nq) return;       // this direction has fewer query tiles (whole CTA leaves)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.y;
  const int q0 = blockIdx.x * kAttQ;
  int* const err = args.err;
  constexpr int kPBufs = P == 1 ? 2 : 1;             // P buffers (the fp16-pair planes leave room for one)
  constexpr int kDepth = kPBufs - 1;                 // tiles by which the P V read-back trails the soft-max
  // shared-memory plan: Q | 2 x (K, V^T) stages | P buffers | barriers | row-maximum exchange
  constexpr uint32_t stage_bytes = (uint32_t)P * (kKBytes + kVBytes);
  uint8_t* sQ = smem;
  uint8_t* sKV = sQ + (size_t)P * kQBytes;
  uint8_t* sP = sKV + 2 * (size_t)stage_bytes;
  uint64_t* bars = (uint64_t*)(sP + (size_t)kPBufs * P * kPBytes);
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t q_full = bar0, kv_full0 = bar0 + 8, kv_empty0 = bar0 + 24, s_full = bar0 + 40, s_free = bar0 + 48,
                 p_full0 = bar0 + 56, pv_full0 = bar0 + 72;
  uint32_t* tmem_slot = (uint32_t*)(bars + 12);
  float* xch = (float*)(bars + 14);                  // [2 parities][2 halves][128 rows]
  const int ntiles = (a.nk + kAttK - 1) / kAttK;

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(kv_full0 + 8 * s, 1); mbar_init(kv_empty0 + 8 * s, 1);
      mbar_init(p_full0 + 8 * s, 256); mbar_init(pv_full0 + 8 * s, 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_free, 256);
    mbar_fence_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ0); tma_prefetch_desc(&tmK0); tma_prefetch_desc(&tmV0);
    if (P == 2) { tma_prefetch_desc(&tmQ1); tma_prefetch_desc(&tmK1); tma_prefetch_desc(&tmV1); }
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_s = tmem_base, tmem_pv = tmem_base + 128;      // S: 128 columns; P V: 2 x 32 columns

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, (uint32_t)P * kQBytes);
      tma_load_3d(smem_u32(sQ), &tmQ0, q_full, 0, a.q_row0 + q0, head);
      if (P == 2) tma_load_3d(smem_u32(sQ) + kQBytes, &tmQ1, q_full, 0, a.q_row0 + q0, head);
      for (int t = 0; t < ntiles; ++t) {
        const int s = t & 1;
        mbar_wait(kv_empty0 + 8 * s, (((uint32_t)t >> 1) & 1u) ^ 1u, err, 41);
        const uint32_t fb = kv_full0 + 8 * s;
        mbar_expect_tx(fb, stage_bytes);
        const uint32_t sk = smem_u32(sKV + (size_t)s * stage_bytes);
        const uint32_t sv = sk + (uint32_t)P * kKBytes;
        const int key0 = a.k_row0 + t * kAttK;
#pragma unroll
        for (int p = 0; p < P; ++p) {
          tma_load_3d(sk + p * kKBytes, p ? &tmK1 : &tmK0, fb, 0, key0, head);
          tma_load_3d(sv + p * kVBytes, p ? &tmV1 : &tmV0, fb, key0, 0, head);
          tma_load_3d(sv + p * kVBytes + kVBytes / 2, p ? &tmV1 : &tmV0, fb, key0 + 64, 0, head);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc_16(kAttQ, kAttK, P == 1);
      const uint32_t idesc_pv = umma_idesc_16(kAttQ, 32, P == 1);
      auto issue_s = [&](int t) {
        const int s = t & 1;
        mbar_wait(kv_full0 + 8 * s, ((uint32_t)t >> 1) & 1u, err, 42);
        if (t > 0) mbar_wait(s_free, ((uint32_t)(t - 1)) & 1u, err, 43);    // soft-max has read S of tile t - 1
        tc_fence_after();
        const uint32_t sk = smem_u32(sKV + (size_t)s * stage_bytes);
        const uint64_t dq0 = umma_desc_sw128(smem_u32(sQ)), dk0 = umma_desc_sw128(sk);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t koff = (uint64_t)(k * 2);
          umma_f16(tmem_s, dq0 + koff, dk0 + koff, idesc_s, k ? 1u : 0u);
          if (P == 2) {
            const uint64_t dq1 = umma_desc_sw128(smem_u32(sQ) + kQBytes), dk1 = umma_desc_sw128(sk + kKBytes);
            umma_f16(tmem_s, dq0 + koff, dk1 + koff, idesc_s, 1u);
            umma_f16(tmem_s, dq1 + koff, dk0 + koff, idesc_s, 1u);
          }
        }
        umma_commit(s_full);
      };
      mbar_wait(q_full, 0, err, 44);
      issue_s(0);
      for (int t = 0; t < ntiles; ++t) {
        if (t + 1 < ntiles) issue_s(t + 1);
        const int s = t & 1;
        // P(t) is in shared memory; its arrival also tells that the P V accumulator of tile t - 2 has been read
        mbar_wait(p_full0 + 8 * s, ((uint32_t)t >> 1) & 1u, err, 45);
        tc_fence_after();
        const uint32_t sv = smem_u32(sKV + (size_t)s * stage_bytes) + (uint32_t)P * kKBytes;
        const uint32_t sp = smem_u32(sP) + (uint32_t)(t % kPBufs) * (uint32_t)P * kPBytes;
        const uint32_t tmem_o = tmem_pv + (uint32_t)s * 32u;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t sub = (uint32_t)(k >> 2);
          const uint64_t koff = (uint64_t)((k & 3) * 2);
          const uint64_t dp0 = umma_desc_sw128(sp + sub * (kPBytes / 2)) + koff;
          const uint64_t dv0 = umma_desc_sw128(sv + sub * (kVBytes / 2)) + koff;
          umma_f16(tmem_o, dp0, dv0, idesc_pv, k ? 1u : 0u);
          if (P == 2) {
            const uint64_t dp1 = umma_desc_sw128(sp + kPBytes + sub * (kPBytes / 2)) + koff;
            const uint64_t dv1 = umma_desc_sw128(sv + kVBytes + sub * (kVBytes / 2)) + koff;
            umma_f16(tmem_o, dp0, dv1, idesc_pv, 1u);
            umma_f16(tmem_o, dp1, dv0, idesc_pv, 1u);
          }
        }
        umma_commit(pv_full0 + 8 * s);
        umma_commit(kv_empty0 + 8 * s);
      }
    }
  } else {
    // ------------------------------- soft-max / accumulate: two threads per query row ------------
    const int qd = warp & 3;                         // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;                // keys [64 half, 64 half + 64) of a tile, output columns [16 half, +16)
    const int row = qd * 32 + lane;
    const uint32_t lane_addr = ((uint32_t)(qd * 32)) << 16;
    float o[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) o[j] = 0.f;
    float m_run = -INFINITY, l_run = 0.f, corr_prev = 0.f;
    // o <- o * corr(tt) + (P V)(tt), this thread's 16 columns
    auto fold_pv = [&](int tt, float c) {
      mbar_wait(pv_full0 + 8 * (tt & 1), ((uint32_t)tt >> 1) & 1u, err, 47);
      tc_fence_after();
      uint32_t v[16];
      tmem_ld_32x32b_x16(tmem_pv + lane_addr + (uint32_t)((tt & 1) * 32 + half * 16), v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) o[j] = fmaf(o[j], c, __uint_as_float(v[j]));
      tc_fence_before();
    };
    for (int t = 0; t < ntiles; ++t) {
      mbar_wait(s_full, (uint32_t)t & 1u, err, 46);
      tc_fence_after();
      float sv_[64];
      {
        uint32_t v0[32], v1[32];                         // both loads in flight, one wait
        tmem_ld_32x32b_x32(tmem_s + lane_addr + (uint32_t)(half * 64), v0);
        tmem_ld_32x32b_x32(tmem_s + lane_addr + (uint32_t)(half * 64 + 32), v1);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) { sv_[j] = __uint_as_float(v0[j]); sv_[32 + j] = __uint_as_float(v1[j]); }
      }
      tc_fence_before();
      mbar_arrive(s_free);
      const int valid = min(64, max(0, a.nk - t * kAttK - half * 64));      // of this thread's 64 keys
      float mloc = -INFINITY;
      if (valid == 64) {
        // four independent chains (a single running maximum is a 64-deep dependency chain for a thread that
        // shares its scheduler with one other warp)
        float m4[4] = {sv_[0], sv_[1], sv_[2], sv_[3]};
#pragma unroll
        for (int j = 4; j < 64; j += 4) {
          m4[0] = fmaxf(m4[0], sv_[j]); m4[1] = fmaxf(m4[1], sv_[j + 1]);
          m4[2] = fmaxf(m4[2], sv_[j + 2]); m4[3] = fmaxf(m4[3], sv_[j + 3]);
        }
        mloc = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      } else {
#pragma unroll
        for (int j = 0; j < 64; ++j)
          if (j < valid) mloc = fmaxf(mloc, sv_[j]);
      }
      float* xc = xch + (t & 1) * 256;
      xc[half * 128 + row] = mloc;
      named_bar_sync(1 + qd, 64);                           // the two warps of this lane quarter
      const float mx = fmaxf(m_run, fmaxf(mloc, xc[(half ^ 1) * 128 + row]));
      const float corr = ex2_fast(m_run - mx);              // first tile: 2^(-inf) = 0
      m_run = mx;
      // this thread's half of the row of P, in the swizzled K-major layout the tensor core reads:
      // sub-tile (= half) / row / 16-byte chunk ^ (row & 7)
      const uint32_t prow = smem_u32(sP) + (uint32_t)(t % kPBufs) * (uint32_t)P * kPBytes +
                            (uint32_t)half * (kPBytes / 2) + (uint32_t)row * 128u;
      float lsum;
      if (valid == 64) lsum = softmax_half_row<P, true>(sv_, mx, 64, prow, row);
      else lsum = softmax_half_row<P, false>(sv_, mx, valid, prow, row);
      l_run = fmaf(l_run, corr, lsum);
      fence_proxy_async();                                  // generic-proxy writes -> visible to the tensor core
      mbar_arrive(p_full0 + 8 * (t & 1));
      if (kDepth == 0) {
        fold_pv(t, corr);
      } else {
        if (t > 0) fold_pv(t - 1, corr_prev);
        corr_prev = corr;
      }
    }
    if (kDepth == 1) fold_pv(ntiles - 1, corr_prev);
    // the row sum is the sum of the two halves
    float* xc = xch + (ntiles & 1) * 256;
    xc[half * 128 + row] = l_run;
    named_bar_sync(1 + qd, 64);
    const float l_tot = l_run + xc[(half ^ 1) * 128 + row];
    const int qi = q0 + row;
    if (qi < a.nq) {
      const float inv = 1.f / l_tot;
      const long long off = (long long)(a.out_row0 + qi) * args.ld_out + head * 32 + half * 16;
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        const float r0 = o[j] * inv, r1 = o[j + 1] * inv, r2 = o[j + 2] * inv, r3 = o[j + 3] * inv;
        if (args.out) *(float4*)(args.out + off + j) = make_float4(r0, r1, r2, r3);
        if (args.out_hi) {
          plane_t h0, l0, h1, l1, h2, l2, h3, l3;
          const bool pr = args.out_lo != nullptr;
          split16(r0, pr, h0, l0); split16(r1, pr, h1, l1); split16(r2, pr, h2, l2); split16(r3, pr, h3, l3);
          *(uint2*)(args.out_hi + off + j) = make_uint2(pack16x2(h0, h1), pack16x2(h2, h3));
          if (pr) *(uint2*)(args.out_lo + off + j) = make_uint2(pack16x2(l0, l1), pack16x2(l2, l3));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// ------------------------------------------------------------------------------------------
// bf16, TWO query tiles per CTA (256 queries), ping-pong: the same 8 soft-max warps alternate between tile A and
// tile B, so while they exponentiate one tile the tensor core runs S = Q'K^T of the next key tile and P V of the
// previous one for the other - the MMA / barrier round trip that a single tile has to wait out (2900 clk per tile
// against 1024 clk of MUFU work, ncu round 2) is covered by useful work, and K / V^T tiles are fetched once for 256
// queries.  One P buffer and one P V accumulator per query tile suffice: P_X(t + 1) is written after the fold of
// (P V)_X(t), which happens a whole other-tile soft-max after its MMA was issued.
//   TMEM: S_A, S_B (2 x 128 columns) + (P V)_A, (P V)_B (2 x 32)        smem: Q 32 KB, K / V^T 2 x 24 KB, P 64 KB
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kAttThreads, 1)
att_fwd2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttArgs args) {
  const AttProblem a = args.pr[blockIdx.z];
  if ((int)blockIdx.x * 2 * kAttQ >= a.nq) return;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.y;
  const int q0 = blockIdx.x * 2 * kAttQ;
  int* const err = args.err;
  constexpr uint32_t stage_bytes = kKBytes + kVBytes;
  uint8_t* sQ = smem;                                  // [2 tiles][128][64]
  uint8_t* sKV = sQ + 2 * (size_t)kQBytes;
  uint8_t* sP = sKV + 2 * (size_t)stage_bytes;         // [2 tiles][32 KB]
  uint64_t* bars = (uint64_t*)(sP + 2 * (size_t)kPBytes);
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t q_full = bar0, kv_full0 = bar0 + 8, kv_empty0 = bar0 + 24, s_full0 = bar0 + 40, s_free0 = bar0 + 56,
                 p_full0 = bar0 + 72, pv_full0 = bar0 + 88;                  // [2] each, indexed by query tile
  uint32_t* tmem_slot = (uint32_t*)(bars + 14);
  float* xch = (float*)(bars + 16);                    // [2 tiles][2 parities][2 halves][128 rows]
  const int ntiles = (a.nk + kAttK - 1) / kAttK;

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(kv_full0 + 8 * s, 1); mbar_init(kv_empty0 + 8 * s, 1);
      mbar_init(s_full0 + 8 * s, 1); mbar_init(s_free0 + 8 * s, 256);
      mbar_init(p_full0 + 8 * s, 256); mbar_init(pv_full0 + 8 * s, 1);
    }
    mbar_fence_init();
  }
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_s = tmem_base, tmem_pv = tmem_base + 256;       // S_A, S_B: 2 x 128 columns; P V: 2 x 32

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, 2 * kQBytes);
      tma_load_3d(smem_u32(sQ), &tmQ, q_full, 0, a.q_row0 + q0, head);
      tma_load_3d(smem_u32(sQ) + kQBytes, &tmQ, q_full, 0, a.q_row0 + q0 + kAttQ, head);
      for (int t = 0; t < ntiles; ++t) {
        const int s = t & 1;
        mbar_wait(kv_empty0 + 8 * s, (((uint32_t)t >> 1) & 1u) ^ 1u, err, 41);
        const uint32_t fb = kv_full0 + 8 * s;
        mbar_expect_tx(fb, stage_bytes);
        const uint32_t sk = smem_u32(sKV + (size_t)s * stage_bytes);
        const uint32_t sv = sk + kKBytes;
        const int key0 = a.k_row0 + t * kAttK;
        tma_load_3d(sk, &tmK, fb, 0, key0, head);
        tma_load_3d(sv, &tmV, fb, key0, 0, head);
        tma_load_3d(sv + kVBytes / 2, &tmV, fb, key0 + 64, 0, head);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc_16(kAttQ, kAttK, true);
      const uint32_t idesc_pv = umma_idesc_16(kAttQ, 32, true);
      auto issue_s = [&](int x, int t) {               // S_x(t) = Q'_x K(t)^T
        const int s = t & 1;
        if (x == 0) mbar_wait(kv_full0 + 8 * s, ((uint32_t)t >> 1) & 1u, err, 42);
        if (t > 0) mbar_wait(s_free0 + 8 * x, ((uint32_t)(t - 1)) & 1u, err, 43);
        tc_fence_after();
        const uint64_t dq = umma_desc_sw128(smem_u32(sQ) + (uint32_t)x * kQBytes);
        const uint64_t dk = umma_desc_sw128(smem_u32(sKV + (size_t)s * stage_bytes));
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(tmem_s + (uint32_t)x * 128u, dq + (uint64_t)(k * 2), dk + (uint64_t)(k * 2), idesc_s, k ? 1u : 0u);
        umma_commit(s_full0 + 8 * x);
      };
      auto issue_pv = [&](int x, int t) {              // (P V)_x(t)
        const int s = t & 1;
        mbar_wait(p_full0 + 8 * x, (uint32_t)t & 1u, err, 45);
        tc_fence_after();
        const uint32_t sv = smem_u32(sKV + (size_t)s * stage_bytes) + kKBytes;
        const uint32_t sp = smem_u32(sP) + (uint32_t)x * kPBytes;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t sub = (uint32_t)(k >> 2);
          const uint64_t koff = (uint64_t)((k & 3) * 2);
          umma_f16(tmem_pv + (uint32_t)x * 32u, umma_desc_sw128(sp + sub * (kPBytes / 2)) + koff,
                   umma_desc_sw128(sv + sub * (kVBytes / 2)) + koff, idesc_pv, k ? 1u : 0u);
        }
        umma_commit(pv_full0 + 8 * x);
      };
      mbar_wait(q_full, 0, err, 44);
      issue_s(0, 0);
      issue_s(1, 0);
      for (int t = 0; t < ntiles; ++t) {
        issue_pv(0, t);
        if (t + 1 < ntiles) issue_s(0, t + 1);
        issue_pv(1, t);
        umma_commit(kv_empty0 + 8 * (t & 1));          // K(t) / V(t) are free once everything issued so far retires
        if (t + 1 < ntiles) issue_s(1, t + 1);
      }
    }
  } else {
    const int qd = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = qd * 32 + lane;
    const uint32_t lane_addr = ((uint32_t)(qd * 32)) << 16;
    float o[2][16];
#pragma unroll
    for (int x = 0; x < 2; ++x)
#pragma unroll
      for (int j = 0; j < 16; ++j) o[x][j] = 0.f;
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f}, corr_prev[2] = {0.f, 0.f};
    auto fold_pv = [&](int x, int tt, float c) {        // o_x <- o_x corr(tt) + (P V)_x(tt)
      mbar_wait(pv_full0 + 8 * x, (uint32_t)tt & 1u, err, 47);
      tc_fence_after();
      uint32_t v[16];
      tmem_ld_32x32b_x16(tmem_pv + lane_addr + (uint32_t)(x * 32 + half * 16), v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) o[x][j] = fmaf(o[x][j], c, __uint_as_float(v[j]));
      tc_fence_before();
    };
    for (int t = 0; t < ntiles; ++t) {
      const int valid = min(64, max(0, a.nk - t * kAttK - half * 64));
#pragma unroll
      for (int x = 0; x < 2; ++x) {
        mbar_wait(s_full0 + 8 * x, (uint32_t)t & 1u, err, 46);
        tc_fence_after();
        float sv_[64];
        {
          uint32_t v0[32], v1[32];
          tmem_ld_32x32b_x32(tmem_s + lane_addr + (uint32_t)(x * 128 + half * 64), v0);
          tmem_ld_32x32b_x32(tmem_s + lane_addr + (uint32_t)(x * 128 + half * 64 + 32), v1);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) { sv_[j] = __uint_as_float(v0[j]); sv_[32 + j] = __uint_as_float(v1[j]); }
        }
        tc_fence_before();
        mbar_arrive(s_free0 + 8 * x);
        float mloc = -INFINITY;
        if (valid == 64) {
          float m4[4] = {sv_[0], sv_[1], sv_[2], sv_[3]};
#pragma unroll
          for (int j = 4; j < 64; j += 4) {
            m4[0] = fmaxf(m4[0], sv_[j]); m4[1] = fmaxf(m4[1], sv_[j + 1]);
            m4[2] = fmaxf(m4[2], sv_[j + 2]); m4[3] = fmaxf(m4[3], sv_[j + 3]);
          }
          mloc = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
        } else {
#pragma unroll
          for (int j = 0; j < 64; ++j)
            if (j < valid) mloc = fmaxf(mloc, sv_[j]);
        }
        float* xc = xch + (x * 2 + (t & 1)) * 256;
        xc[half * 128 + row] = mloc;
        named_bar_sync(1 + qd, 64);
        const float mx = fmaxf(m_run[x], fmaxf(mloc, xc[(half ^ 1) * 128 + row]));
        const float corr = ex2_fast(m_run[x] - mx);
        m_run[x] = mx;
        // (P V)_x(t - 1) was issued a whole other-tile soft-max ago: fold it, which also frees this tile's P buffer
        if (t > 0) fold_pv(x, t - 1, corr_prev[x]);
        corr_prev[x] = corr;
        const uint32_t prow = smem_u32(sP) + (uint32_t)x * kPBytes + (uint32_t)half * (kPBytes / 2) + (uint32_t)row * 128u;
        float lsum;
        if (valid == 64) lsum = softmax_half_row<1, true>(sv_, mx, 64, prow, row);
        else lsum = softmax_half_row<1, false>(sv_, mx, valid, prow, row);
        l_run[x] = fmaf(l_run[x], corr, lsum);
        fence_proxy_async();
        mbar_arrive(p_full0 + 8 * x);
      }
    }
#pragma unroll
    for (int x = 0; x < 2; ++x) {
      fold_pv(x, ntiles - 1, corr_prev[x]);
      float* xc = xch + (x * 2 + (ntiles & 1)) * 256;
      xc[half * 128 + row] = l_run[x];
      named_bar_sync(1 + qd, 64);
      const float l_tot = l_run[x] + xc[(half ^ 1) * 128 + row];
      const int qi = q0 + x * kAttQ + row;
      if (qi < a.nq) {
        const float inv = 1.f / l_tot;
        const long long off = (long long)(a.out_row0 + qi) * args.ld_out + head * 32 + half * 16;
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const float r0 = o[x][j] * inv, r1 = o[x][j + 1] * inv, r2 = o[x][j + 2] * inv, r3 = o[x][j + 3] * inv;
          if (args.out) *(float4*)(args.out + off + j) = make_float4(r0, r1, r2, r3);
          if (args.out_hi) {
            const __nv_bfloat162 a2 = __floats2bfloat162_rn(r0, r1), b2 = __floats2bfloat162_rn(r2, r3);
            *(uint2*)(args.out_hi + off + j) = make_uint2(*(const uint32_t*)&a2, *(const uint32_t*)&b2);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// q / k / v: fp32 rows [n][ld] (columns head * 32 + d).  Qp, Kp: [heads][n_pad][64]; Vt: [heads][32][n_pad].
// The rows form two segments (source cloud [0, split), target cloud [split, n)); the second one starts at the
// 128-aligned packed row split_pad: TMA needs the innermost start of a box 16-byte aligned, and tokens are the
// innermost axis of V^T.
__global__ void att_pack_kernel(const float* __restrict__ q, int ldq, const float* __restrict__ k, int ldk,
                                const float* __restrict__ v, int ldv, int n, int split, int split_pad, int n_pad,
                                int heads, float q_scale, plane_t* __restrict__ qp_hi, plane_t* __restrict__ qp_lo,
                                plane_t* __restrict__ kp_hi, plane_t* __restrict__ kp_lo, plane_t* __restrict__ vt_hi,
                                plane_t* __restrict__ vt_lo) {
  const bool pair = qp_lo != nullptr;
  const long long total = (long long)heads * n_pad * 64;
  const long long stride = (long long)gridDim.x * blockDim.x;
  auto source_row = [&](int ptok) -> int {      // packed row -> input row, -1 for padding
    if (ptok < split) return ptok;
    const int r = ptok - split_pad + split;
    return (ptok >= split_pad && r < n) ? r : -1;
  };
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int d = (int)(i & 63);
    const long long r = i >> 6;
    const int tok = source_row((int)(r % n_pad)), h = (int)(r / n_pad);
    float qv = 0.f, kv = 0.f;
    if (d < 32 && tok >= 0) {
      qv = q[(long long)tok * ldq + h * 32 + d] * q_scale;
      kv = k[(long long)tok * ldk + h * 32 + d];
    }
    plane_t a, b;
    split16(qv, pair, a, b);
    qp_hi[i] = a;
    if (pair) qp_lo[i] = b;
    split16(kv, pair, a, b);
    kp_hi[i] = a;
    if (pair) kp_lo[i] = b;
  }
  // V transposed: thread per (head, d, token) with token fastest for coalesced writes
  const long long totv = (long long)heads * 32 * n_pad;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < totv; i += stride) {
    const int tok = source_row((int)(i % n_pad));
    const long long r = i / n_pad;
    const int d = (int)(r & 31), h = (int)(r >> 5);
    const float vv = tok >= 0 ? v[(long long)tok * ldv + h * 32 + d] : 0.f;
    plane_t a, b;
    split16(vv, pair, a, b);
    vt_hi[i] = a;
    if (pair) vt_lo[i] = b;
  }
}

static inline int att_pad128(int n) { return ((n + 127) / 128) * 128; }
static inline int att_n_pad(int n) { return att_pad128(n) + 256; }      // two padded segments + one spare tile
static size_t att_plane_elems(int n_pad, int heads) { return (size_t)heads * n_pad * 64; }

extern "C" size_t drb_mha_tc_workspace_bytes(int n, int heads, int planes) {
  const int n_pad = att_n_pad(n);
  // Qp + Kp ([heads][n_pad][64]) + Vt ([heads][32][n_pad]) per plane
  return (size_t)planes * (2 * att_plane_elems(n_pad, heads) + (size_t)heads * 32 * n_pad) * sizeof(plane_t) + 1024;
}

extern "C" int drb_mha_tc_pack(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv, int n,
                               int split, int heads, int planes, float scale, void* workspace, size_t workspace_bytes,
                               cudaStream_t stream) {
  DRB_REQUIRE(q && k && v && workspace && n > 0 && heads > 0 && (planes == 1 || planes == 2), "drb_mha_tc_pack: bad arguments");
  DRB_REQUIRE(split >= 0 && split <= n, "drb_mha_tc_pack: split outside [0, n]");
  DRB_REQUIRE(workspace_bytes >= drb_mha_tc_workspace_bytes(n, heads, planes), "drb_mha_tc_pack: workspace too small");
  DRB_REQUIRE(((uintptr_t)workspace & 1023) == 0, "drb_mha_tc_pack: workspace must be 1024-byte aligned");
  const int n_pad = att_n_pad(n);
  const size_t pe = att_plane_elems(n_pad, heads), ve = (size_t)heads * 32 * n_pad;
  plane_t* base = (plane_t*)workspace;
  plane_t* qp_hi = base; plane_t* kp_hi = qp_hi + pe; plane_t* vt_hi = kp_hi + pe;
  plane_t *qp_lo = nullptr, *kp_lo = nullptr, *vt_lo = nullptr;
  if (planes == 2) { qp_lo = vt_hi + ve; kp_lo = qp_lo + pe; vt_lo = kp_lo + pe; }
  const long long total = (long long)heads * n_pad * 64;
  int grid = (int)((total + 255) / 256);
  if (grid > 148 * 16) grid = 148 * 16;
  att_pack_kernel<<<grid, 256, 0, stream>>>(q, ldq, k, ldk, v, ldv, n, split, att_pad128(split), n_pad, heads,
                                            scale * 1.4426950408889634f, qp_hi, qp_lo, kp_hi, kp_lo, vt_hi, vt_lo);
  DRB_LAUNCH_OK();
  return 0;
}

// Attention of the queries of segment q_seg (0: rows [0, split), 1: rows [split, n), -1: both segments in one
// launch) against the keys / values of ONE packed workspace: cross == 0: the queries' own segment (self-attention),
// cross != 0: the other segment.  Output rows are the queries' input row indices.
extern "C" int drb_mha_tc_forward(const void* workspace, int n, int split, int heads, int planes, int q_seg, int cross,
                                  float* out, void* out_hi, void* out_lo, int ld_out, cudaStream_t stream) {
  DRB_REQUIRE(workspace && n > 0 && heads > 0 && (planes == 1 || planes == 2), "drb_mha_tc_forward: bad arguments");
  DRB_REQUIRE(split >= 0 && split <= n && q_seg >= -1 && q_seg <= 1, "drb_mha_tc_forward: bad segment");
  DRB_REQUIRE((out || out_hi) && ld_out % 8 == 0, "drb_mha_tc_forward: bad output");
  const int split_pad = att_pad128(split);
  const int n_pad = att_n_pad(n);
  AttArgs a;
  memset(&a, 0, sizeof(a));
  int nprob = 0, max_q = 0;
  for (int seg = 0; seg < 2; ++seg) {
    if (q_seg >= 0 && q_seg != seg) continue;
    const int kseg = cross ? 1 - seg : seg;
    AttProblem& p = a.pr[nprob];
    p.nq = seg ? n - split : split;
    p.nk = kseg ? n - split : split;
    p.q_row0 = seg ? split_pad : 0;
    p.k_row0 = kseg ? split_pad : 0;
    p.out_row0 = seg ? split : 0;
    if (p.nq == 0) continue;
    DRB_REQUIRE(p.nk > 0, "drb_mha_tc_forward: empty key segment");
    if (p.nq > max_q) max_q = p.nq;
    ++nprob;
  }
  if (nprob == 0) return 0;
  const size_t pe = att_plane_elems(n_pad, heads), ve = (size_t)heads * 32 * n_pad;
  const plane_t* base = (const plane_t*)workspace;
  const plane_t* qp[2] = {base, nullptr};
  const plane_t* kp[2] = {base + pe, nullptr};
  const plane_t* vt[2] = {base + 2 * pe, nullptr};
  if (planes == 2) { qp[1] = vt[0] + ve; kp[1] = qp[1] + pe; vt[1] = kp[1] + pe; }
  CUtensorMap mQ[2], mK[2], mV[2];
  memset(mQ, 0, sizeof(mQ)); memset(mK, 0, sizeof(mK)); memset(mV, 0, sizeof(mV));
  const uint64_t rdims[3] = {64, (uint64_t)n_pad, (uint64_t)heads};
  const uint64_t rstr[2] = {64 * 2, (uint64_t)n_pad * 64 * 2};
  const uint32_t rbox[3] = {64, 128, 1};
  const uint64_t vdims[3] = {(uint64_t)n_pad, 32, (uint64_t)heads};
  const uint64_t vstr[2] = {(uint64_t)n_pad * 2, (uint64_t)n_pad * 32 * 2};
  const uint32_t vbox[3] = {64, 32, 1};
  int rc;
  for (int p = 0; p < planes; ++p) {
    const bool bf = planes == 1;
    if ((rc = igemm_make_map(&mQ[p], qp[p], 3, rdims, rstr, rbox, bf))) return rc;
    if ((rc = igemm_make_map(&mK[p], kp[p], 3, rdims, rstr, rbox, bf))) return rc;
    if ((rc = igemm_make_map(&mV[p], vt[p], 3, vdims, vstr, vbox, bf))) return rc;
  }
  if (planes == 1) { mQ[1] = mQ[0]; mK[1] = mK[0]; mV[1] = mV[0]; }
  a.heads = heads; a.planes = planes;
  a.out = out; a.out_hi = (plane_t*)out_hi; a.out_lo = (plane_t*)out_lo; a.ld_out = ld_out;
  a.err = igemm_err_flag();
  const int pbufs = planes == 1 ? 2 : 1;
  const size_t smem = 1024 + (size_t)planes * (kQBytes + 2 * (kKBytes + kVBytes) + (size_t)pbufs * kPBytes) + 14 * 8 +
                      2 * 2 * 128 * sizeof(float) + 16;
  dim3 grid((unsigned)cdiv(max_q, kAttQ), (unsigned)heads, (unsigned)nprob);
  static int two_tiles = -1;     // DRB_ATT_TWO_TILES=0: the one-tile kernel for bf16 too
  if (two_tiles < 0) { const char* env = getenv("DRB_ATT_TWO_TILES"); two_tiles = env ? atoi(env) : 1; }
  if (planes == 1 && two_tiles && max_q > kAttQ && a.out_lo == nullptr) {
    const size_t smem2 = 1024 + 2 * (size_t)kQBytes + 2 * (size_t)(kKBytes + kVBytes) + 2 * (size_t)kPBytes + 16 * 8 +
                         2 * 2 * 2 * 128 * sizeof(float) + 16;
    DRB_CUDA_OK(cudaFuncSetAttribute(att_fwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    dim3 grid2((unsigned)cdiv(max_q, 2 * kAttQ), (unsigned)heads, (unsigned)nprob);
    att_fwd2_kernel<<<grid2, kAttThreads, smem2, stream>>>(mQ[0], mK[0], mV[0], a);
    DRB_LAUNCH_OK();
    return 0;
  }
  if (planes == 1) {
    DRB_CUDA_OK(cudaFuncSetAttribute(att_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    att_fwd_kernel<1><<<grid, kAttThreads, smem, stream>>>(mQ[0], mQ[1], mK[0], mK[1], mV[0], mV[1], a);
  } else {
    DRB_CUDA_OK(cudaFuncSetAttribute(att_fwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    att_fwd_kernel<2><<<grid, kAttThreads, smem, stream>>>(mQ[0], mQ[1], mK[0], mK[1], mV[0], mV[1], a);
  }
  DRB_LAUNCH_OK();
  return 0;
}

}  // namespace drb
