// Implicit-GEMM 3-D convolution / linear layer on the 5th-gen tensor cores (tcgen05 + TMEM + TMA).
//
// Replaces every nn.Conv3d (conerf/model/resnet3d.py:81-86,120; feature_pyramid_net.py:24,33) and
// every nn.Linear / in_proj (conerf/register/transformer.py:128-138, nerf_regtr.py:268-270) on the
// registration path.  One persistent kernel, warp specialised:
//   warp 0   : TMA producer  - 5-D tiled loads of channels-last activation boxes (zero fill out of
//              bounds == convolution padding) and 3-D loads of [tap][Cout][Cin] weight tiles
//   warp 1   : TMEM allocator + single-thread tcgen05.mma issuer (128 x BN x 16 MMAs)
//   warps 2-9: accumulate + epilogue - drain each finished K chunk from TMEM (tcgen05.ld) into fp32
//              registers, then fused bias / scale / residual / ReLU and stores of fp32 and/or
//              16-bit planes for the next layer
// Precision: operands are 16-bit "planes".  planes == 1: plain bf16.  planes == 2: every fp32 value x
// is carried as hi = fp16(x), lo = fp16(x - hi) (22 significand bits) and the product is accumulated
// as Ah*Bh + Ah*Bl + Al*Bh in fp32 (error ~2^-22 per product, i.e. fp32-grade results from the
// 16-bit tensor pipe at 3 MMAs per product instead of the 2x slower, 2^-11 accurate TF32 pipe).
// Weights are pre-scaled by a power of two into fp16's normal range; acc_scale undoes it.
#include "common.cuh"

#include <cudaTypedefs.h>
#include <stdlib.h>

namespace drb {

static constexpr int kBM = 128;       // accumulator rows = TMEM lanes
static constexpr int kBK = 64;        // K chunk: 64 x 16-bit = one 128-byte swizzle row
static constexpr int kMaxStages = 8;
static constexpr int kAccWarps = 8;   // accumulate / epilogue warps
static constexpr int kThreads = 64 + kAccWarps * 32;   // warp 0 TMA, warp 1 MMA, warps 2-9 accumulate

struct IgemmArgs {
  int G, D, H, W;        // output (== input) spatial extent, stride-1 "same" convolution
  int Cin, Cout;
  int kd, kh, kw;        // kernel extent
  int pd, ph, pw;        // padding
  int bg, bd, bh, bw;    // spatial box of one 128-row tile
  int BN;                // N tile: 64, 128, 192 or 256
  int planes;            // 1 or 2
  int stages;
  int chunk;             // k-iterations accumulated in TMEM before the fp32 register add
  int cs;                // cluster size along M (1, 2 or 4): the weight tile is TMA-multicast
  int splits;            // split-K factor (> 1: every K slice stores its partial tile into `splitk_ws`, a second
                         // kernel adds the slices in a fixed order: run-to-run bit-stable, no atomics)
  float* splitk_ws;      // [splits][M][ld] fp32
  long long m_total;     // M = G*D*H*W
  int kper;              // k-iterations per split (multiple of chunk)
  double* bn_accum;      // optional [G][Cout][2]: per-channel sum / sum of squares of the output (fused BN stats)
  const int* tile_list;  // optional: compacted list of M-tile indices to compute (output-sparse conv)
  const int* tile_count; // device scalar: number of entries in tile_list
  int relu;
  float acc_scale;         // multiplies the raw accumulator (undoes the weight pre-scale)
  const float* acc_scale_dev0;   // optional device scalars multiplied into acc_scale (gradient / weight
  const float* acc_scale_dev1;   // pre-scales that only exist on the device)
  float out_scale;
  const float* bias;       // [Cout] or null
  const float* residual;   // [M, ld] fp32 or null; out = relu((acc*acc_scale + bias)*out_scale + residual)
  float* out;              // [M, ld] fp32 or null
  plane_t* out_hi;         // [M, ld] or null
  plane_t* out_lo;         // [M, ld] or null (present => fp16 pair mode)
  long long ld;            // row pitch (elements) of out / residual / out_hi / out_lo
  int res_d, res_h, res_w; // > 0: residual is a coarser volume added with nearest x2 up-sampling
  int* err;
};

// Fused epilogue of 32 consecutive output channels of one row.
__device__ __forceinline__ void epilogue_store32(const IgemmArgs& a, float (&f)[32], long long m, int n,
                                                 int split, float acc_scale, long long m_res) {
  const long long off = m * a.ld + n;
  const long long roff = m_res * a.ld + n;     // residual row (== m unless the residual is the coarser FPN level)
  const bool full = (n + 32 <= a.Cout);
#pragma unroll
  for (int j = 0; j < 32; ++j) f[j] *= acc_scale;
  if (a.splits > 1) {
    // split-K: this K slice's partial tile goes to its own plane of the workspace (plain stores)
    float* dst = a.splitk_ws + (long long)split * a.m_total * a.ld + off;
    if (full) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) *(float4*)(dst + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (n + j < a.Cout) dst[j] = f[j];
    }
    return;
  }
  if (a.bias) {
    if (full) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b4 = __ldg((const float4*)(a.bias + n + j));
        f[j] += b4.x; f[j + 1] += b4.y; f[j + 2] += b4.z; f[j + 3] += b4.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (n + j < a.Cout) f[j] += __ldg(a.bias + n + j);
    }
  }
  if (a.out_scale != 1.f) {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] *= a.out_scale;
  }
  if (a.residual) {
    if (full) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 r4 = *(const float4*)(a.residual + roff + j);
        f[j] += r4.x; f[j + 1] += r4.y; f[j + 2] += r4.z; f[j + 3] += r4.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (n + j < a.Cout) f[j] += a.residual[roff + j];
    }
  }
  if (a.relu) {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
  }
  if (a.out) {
    if (full) {
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *(float4*)(a.out + off + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (n + j < a.Cout) a.out[off + j] = f[j];
    }
  }
  if (a.out_hi) {
    const bool pair = a.out_lo != nullptr;
    if (full) {
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        uint32_t hw[4], lw[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          plane_t h0_, l0_, h1_, l1_;
          split16(f[j + 2 * u], pair, h0_, l0_);
          split16(f[j + 2 * u + 1], pair, h1_, l1_);
          hw[u] = pack16x2(h0_, h1_);
          lw[u] = pack16x2(l0_, l1_);
        }
        *(uint4*)(a.out_hi + off + j) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
        if (pair) *(uint4*)(a.out_lo + off + j) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (n + j < a.Cout) {
          plane_t h_, l_;
          split16(f[j], pair, h_, l_);
          a.out_hi[off + j] = h_;
          if (pair) a.out_lo[off + j] = l_;
        }
      }
    }
  }
}

// Fused BatchNorm statistics: column sums of one 32-column block over the warp's 32 rows by a
// transposing butterfly (31 shuffles; lane l ends up owning one column), then fp64 atomics.
__device__ __forceinline__ float warp_colsum32(float (&t)[32], int lane, int& col) {
  col = 0;
#pragma unroll
  for (int W = 16, mask = 16; mask >= 1; W >>= 1, mask >>= 1) {
    const bool upper = (lane & mask) != 0;
#pragma unroll
    for (int j = 0; j < W; ++j) {
      const float send = upper ? t[j] : t[j + W];
      const float keep = upper ? t[j + W] : t[j];
      t[j] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
    }
    if (upper) col += W;
  }
  return t[0];
}

// The tensor core adds every MMA into the fp32 TMEM accumulator with truncation, so a long K chain
// (K = 6912 -> 1296 MMAs in pair mode) drifts by ~1e-5.  The chain is therefore cut into chunks of
// `chunk` k-iterations that ping-pong between two TMEM regions; the 8 accumulate warps drain each
// finished chunk with tcgen05.ld and add it into fp32 registers (round-to-nearest), which also lets
// the MMA warp start the next tile while the previous tile's epilogue is still storing.
template <bool kBnStats>
__global__ void __launch_bounds__(kThreads, 1)
igemm_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
             const __grid_constant__ CUtensorMap tmB0, const __grid_constant__ CUtensorMap tmB1,
             const IgemmArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment is required by SWIZZLE_128B; dynamic smem base is only 16 B aligned by contract.
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const uint32_t a_bytes = kBM * kBK * 2;                 // 16 KB per plane
  const uint32_t b_bytes = (uint32_t)a.BN * kBK * 2;      // BN rows x 128 B per plane
  const uint32_t stage_bytes = (uint32_t)a.planes * (a_bytes + b_bytes);

  uint64_t* bars = (uint64_t*)(smem + (size_t)a.stages * stage_bytes);
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (kMaxStages + s); };
  auto tfull_bar = [&](int s) { return bar0 + 8u * (2 * kMaxStages + s); };
  auto tempty_bar = [&](int s) { return bar0 + 8u * (2 * kMaxStages + 2 + s); };
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * kMaxStages + 4);

  const int tiles_w = (a.W + a.bw - 1) / a.bw;
  const int tiles_h = (a.H + a.bh - 1) / a.bh;
  const int tiles_d = (a.D + a.bd - 1) / a.bd;
  const int tiles_g = (a.G + a.bg - 1) / a.bg;
  const int tiles_n = (a.Cout + a.BN - 1) / a.BN;
  // Clusters of `cs` CTAs take `cs` consecutive M tiles of the same N tile; each CTA loads 1/cs of the
  // weight rows and multicasts them, so the L2 -> SM traffic of B drops by cs.
  const int cs = a.cs;
  const uint32_t crank = cs > 1 ? cluster_ctarank() : 0u;
  const uint16_t cmask = (uint16_t)((1u << cs) - 1u);
  const int tiles_m = tiles_w * tiles_h * tiles_d * tiles_g;
  const int groups_m = (tiles_m + cs - 1) / cs;
  // output-sparse mode: only the listed M tiles are computed (cs == 1, splits == 1 enforced by the host)
  const int listed = a.tile_list ? *a.tile_count : 0;
  const int total_tiles = (a.tile_list ? listed : groups_m) * tiles_n;   // per cluster: cs M tiles
  const int total_items = total_tiles * a.splits;   // (super tile, K slice) work items
  const int cluster_id = blockIdx.x / cs;
  const int num_clusters = gridDim.x / cs;
  const int kchunks = a.Cin / kBK;
  const int taps = a.kd * a.kh * a.kw;
  const int kiters = taps * kchunks;
  const uint32_t tmem_cols = (2 * a.BN <= 128) ? 128 : (2 * a.BN <= 256) ? 256 : 512;

  if (threadIdx.x == 0) {
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), (uint32_t)cs);      // one tcgen05.commit from every CTA of the cluster
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), kAccWarps * 32);
    }
    mbar_fence_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmB0);
    if (a.planes == 2) {
      tma_prefetch_desc(&tmA1);
      tma_prefetch_desc(&tmB1);
    }
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), tmem_cols);
  tc_fence_before();
  __syncthreads();
  if (cs > 1) cluster_sync_all();    // peers' barriers must exist before any multicast / remote commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto decode_tile = [&](int t, int& n0, int& w0, int& h0, int& d0, int& g0) {
    int n = t % tiles_n;
    int m = a.tile_list ? a.tile_list[t / tiles_n] : (t / tiles_n) * cs + (int)crank;
    n0 = n * a.BN;
    if (m >= tiles_m) {              // padding tile of the last group: all rows out of bounds
      w0 = 0; h0 = 0; d0 = 0; g0 = a.G;
      return;
    }
    int tw = m % tiles_w; m /= tiles_w;
    int th = m % tiles_h; m /= tiles_h;
    int td = m % tiles_d; m /= tiles_d;
    w0 = tw * a.bw; h0 = th * a.bh; d0 = td * a.bd; g0 = m * a.bg;
  };

  if (warp == 0) {
    // ------------------------------- TMA producer -------------------------------------------
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int it = cluster_id; it < total_items; it += num_clusters) {
        const int t = it / a.splits, sp = it % a.splits;
        const int ki0 = sp * a.kper, ki1 = min(kiters, ki0 + a.kper);
        int n0, w0, h0, d0, g0;
        decode_tile(t, n0, w0, h0, d0, g0);
        {
          for (int ki = ki0; ki < ki1; ++ki) {
            const int tap = ki / kchunks, kc = ki - tap * kchunks;
            const int tw = tap % a.kw, th = (tap / a.kw) % a.kh, td = tap / (a.kw * a.kh);
            mbar_wait(empty_bar(s), ph ^ 1u, a.err, 1);
            const uint32_t fb = full_bar(s);
            mbar_expect_tx(fb, stage_bytes);
            const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
            const uint32_t sb = sa + (uint32_t)a.planes * a_bytes;
            tma_load_5d(sa, &tmA0, fb, kc * kBK, w0 + tw - a.pw, h0 + th - a.ph, d0 + td - a.pd, g0);
            if (a.planes == 2)
              tma_load_5d(sa + a_bytes, &tmA1, fb, kc * kBK, w0 + tw - a.pw, h0 + th - a.ph,
                          d0 + td - a.pd, g0);
            if (cs == 1) {
              tma_load_3d(sb, &tmB0, fb, kc * kBK, n0, tap);
              if (a.planes == 2) tma_load_3d(sb + b_bytes, &tmB1, fb, kc * kBK, n0, tap);
            } else {
              // this CTA's slice of the weight rows, delivered to every CTA of the cluster
              const int rows = a.BN / cs;
              const uint32_t soff = crank * (uint32_t)rows * (kBK * 2);
              tma_load_3d_mc(sb + soff, &tmB0, fb, kc * kBK, n0 + (int)crank * rows, tap, cmask);
              if (a.planes == 2)
                tma_load_3d_mc(sb + b_bytes + soff, &tmB1, fb, kc * kBK, n0 + (int)crank * rows, tap, cmask);
            }
            if (++s == a.stages) { s = 0; ph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer ----------------------------------------------
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_16(kBM, a.BN, a.planes == 1);
      int s = 0;
      uint32_t ph = 0;
      uint32_t cc = 0;                       // running chunk counter -> TMEM region + phase
      for (int it = cluster_id; it < total_items; it += num_clusters) {
        const int sp = it % a.splits;
        const int ki0 = sp * a.kper, ki1 = min(kiters, ki0 + a.kper);
        for (int k0 = ki0; k0 < ki1; k0 += a.chunk, ++cc) {
          const uint32_t r = cc & 1u, rph = (cc >> 1) & 1u;
          mbar_wait(tempty_bar(r), rph ^ 1u, a.err, 2);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + r * (uint32_t)a.BN;
          const int kend = min(k0 + a.chunk, ki1);
          for (int ki = k0; ki < kend; ++ki) {
            mbar_wait(full_bar(s), ph, a.err, 3);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
            const uint32_t sb = sa + (uint32_t)a.planes * a_bytes;
            const uint64_t da0 = umma_desc_sw128(sa);
            const uint64_t db0 = umma_desc_sw128(sb);
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k) {
              const uint64_t koff = (uint64_t)(k * 2);  // 16 elements = 32 B = 2 x 16 B units
              umma_f16(tmem_d, da0 + koff, db0 + koff, idesc, (ki != k0 || k != 0) ? 1u : 0u);
              if (a.planes == 2) {
                const uint64_t da1 = umma_desc_sw128(sa + a_bytes);
                const uint64_t db1 = umma_desc_sw128(sb + b_bytes);
                umma_f16(tmem_d, da0 + koff, db1 + koff, idesc, 1u);
                umma_f16(tmem_d, da1 + koff, db0 + koff, idesc, 1u);
              }
            }
            // frees the smem stage (in every CTA that multicasts into it) when the MMAs retire
            if (cs == 1) umma_commit(empty_bar(s)); else umma_commit_mc(empty_bar(s), cmask);
            if (++s == a.stages) { s = 0; ph ^= 1u; }
          }
          umma_commit(tfull_bar(r));                  // chunk complete -> accumulate warps
        }
      }
    }
  } else {
    // ------------------------------- accumulate + epilogue ------------------------------------
    const int q = warp & 3;                  // TMEM lane quarter this warp may access
    const int colhalf = (warp - 2) >> 2;     // warps 2-5: first half of the columns, 6-9: second
    const int half = a.BN >> 1;              // columns per warp (multiple of 32)
    const int row = q * 32 + lane;           // accumulator row == tile-local voxel
    float acc_scale = a.acc_scale;
    if (a.acc_scale_dev0) acc_scale *= *a.acc_scale_dev0;
    if (a.acc_scale_dev1) acc_scale *= *a.acc_scale_dev1;
    float acc[128];
    uint32_t cc = 0;
    for (int it = cluster_id; it < total_items; it += num_clusters) {
      const int t = it / a.splits, sp = it % a.splits;
      const int ki0 = sp * a.kper, ki1 = min(kiters, ki0 + a.kper);
      int n0, w0, h0, d0, g0;
      decode_tile(t, n0, w0, h0, d0, g0);
      int rr = row;
      const int ww = w0 + rr % a.bw; rr /= a.bw;
      const int hh = h0 + rr % a.bh; rr /= a.bh;
      const int dd = d0 + rr % a.bd; rr /= a.bd;
      const int gg = g0 + rr;
      const bool row_ok = (ww < a.W) && (hh < a.H) && (dd < a.D) && (gg < a.G);
      const long long m = (((long long)gg * a.D + dd) * a.H + hh) * a.W + ww;
      const long long m_res = a.res_d > 0
                                  ? (((long long)gg * a.res_d + (dd >> 1)) * a.res_h + (hh >> 1)) * a.res_w + (ww >> 1)
                                  : m;
#pragma unroll
      for (int j = 0; j < 128; ++j) acc[j] = 0.f;
      for (int k0 = ki0; k0 < ki1; k0 += a.chunk, ++cc) {
        const uint32_t r = cc & 1u, rph = (cc >> 1) & 1u;
        mbar_wait(tfull_bar(r), rph, a.err, 4);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + r * (uint32_t)a.BN +
                               (uint32_t)(colhalf * half);
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          if (b * 32 < half) {               // warp-uniform
            uint32_t v[32];
            tmem_ld_32x32b_x32(taddr + (uint32_t)(b * 32), v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[b * 32 + j] += __uint_as_float(v[j]);
          }
        }
        tc_fence_before();
        mbar_arrive(tempty_bar(r));
      }
      if (kBnStats) {
        // all 32 lanes take part (shuffles); rows outside the volume contribute zeros
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          if (b * 32 < half) {
            float tsum[32];
            int col;
#pragma unroll
            for (int j = 0; j < 32; ++j) tsum[j] = row_ok ? acc[b * 32 + j] * acc_scale : 0.f;
            const float csum = warp_colsum32(tsum, lane, col);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float v = row_ok ? acc[b * 32 + j] * acc_scale : 0.f;
              tsum[j] = v * v;
            }
            const float csq = warp_colsum32(tsum, lane, col);
            const int cidx = n0 + colhalf * half + b * 32 + col;
            if (cidx < a.Cout && g0 < a.G) {
              double* dst = a.bn_accum + ((long long)g0 * a.Cout + cidx) * 2;
              atomicAdd(dst, (double)csum);
              atomicAdd(dst + 1, (double)csq);
            }
          }
        }
      }
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int n = n0 + colhalf * half + b * 32;
        if (b * 32 < half && row_ok && n < a.Cout) {
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = acc[b * 32 + j];
          epilogue_store32(a, f, m, n, sp, acc_scale, m_res);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (cs > 1) cluster_sync_all();    // no CTA may exit while a peer can still signal its barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------
// bf16, TWO 128-row M tiles per work item against one 256-wide weight tile (igemm2_kernel).
// A 128 x 256 x 64 bf16 k-iteration moves 48 KB from L2 for 4 MMAs of 128 clk: 96 B/clk per SM, and 200 KB of
// shared memory buffer only ~2100 clk of tensor work - about one TMA round trip - so the one-tile kernel reaches
// 42-54 % of the MMA rate on the large convolutions (profiles/r02_igemm_per_shape.txt).  Sharing the weight tile
// between two M tiles cuts the traffic to 64 B/clk per SM and makes the same shared memory buffer 3200 clk of work.
// bf16 operands need no chunked fp32 accumulation (their own rounding is 2^-9), so the two accumulators fill TMEM
// (2 x 256 columns) and the epilogue reads them once per item; the price is that an item's epilogue does not
// overlap the next item's MMAs (a few % at K >= 1728).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 1)
igemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const IgemmArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr uint32_t a_bytes = kBM * kBK * 2;              // 16 KB per M tile
  constexpr uint32_t b_bytes = 256 * kBK * 2;              // 32 KB
  constexpr uint32_t stage_bytes = 2 * a_bytes + b_bytes;  // 64 KB
  uint64_t* bars = (uint64_t*)(smem + (size_t)a.stages * stage_bytes);
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (kMaxStages + s); };
  const uint32_t tfull_bar = bar0 + 8u * (2 * kMaxStages), tempty_bar = bar0 + 8u * (2 * kMaxStages + 1);
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * kMaxStages + 4);

  const int tiles_w = (a.W + a.bw - 1) / a.bw;
  const int tiles_h = (a.H + a.bh - 1) / a.bh;
  const int tiles_d = (a.D + a.bd - 1) / a.bd;
  const int tiles_g = (a.G + a.bg - 1) / a.bg;
  const int tiles_n = (a.Cout + 255) / 256;
  const int tiles_m = tiles_w * tiles_h * tiles_d * tiles_g;
  const int listed = a.tile_list ? *a.tile_count : tiles_m;
  const int pairs_m = (listed + 1) / 2;
  const int total_items = pairs_m * tiles_n;
  const int kchunks = a.Cin / kBK;
  const int taps = a.kd * a.kh * a.kw;
  const int kiters = taps * kchunks;

  if (threadIdx.x == 0) {
    for (int s = 0; s < a.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, kAccWarps * 32);
    mbar_fence_init();
  }
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // item -> N tile and the two M tiles (index `which` of the pair); a missing second tile has all rows out of bounds
  auto decode = [&](int it, int which, int& n0, int& w0, int& h0, int& d0, int& g0) {
    n0 = (it % tiles_n) * 256;
    const int li = (it / tiles_n) * 2 + which;
    int m = li < listed ? (a.tile_list ? a.tile_list[li] : li) : tiles_m;
    if (m >= tiles_m) { w0 = 0; h0 = 0; d0 = 0; g0 = a.G; return; }
    int tw = m % tiles_w; m /= tiles_w;
    int th = m % tiles_h; m /= tiles_h;
    int td = m % tiles_d; m /= tiles_d;
    w0 = tw * a.bw; h0 = th * a.bh; d0 = td * a.bd; g0 = m * a.bg;
  };

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int it = blockIdx.x; it < total_items; it += gridDim.x) {
        int n0, w0, h0, d0, g0, n1, w1, h1, d1, g1;
        decode(it, 0, n0, w0, h0, d0, g0);
        decode(it, 1, n1, w1, h1, d1, g1);
        for (int ki = 0; ki < kiters; ++ki) {
          const int tap = ki / kchunks, kc = ki - tap * kchunks;
          const int tw = tap % a.kw, th = (tap / a.kw) % a.kh, td = tap / (a.kw * a.kh);
          mbar_wait(empty_bar(s), ph ^ 1u, a.err, 1);
          const uint32_t fb = full_bar(s);
          mbar_expect_tx(fb, stage_bytes);
          const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
          tma_load_5d(sa, &tmA, fb, kc * kBK, w0 + tw - a.pw, h0 + th - a.ph, d0 + td - a.pd, g0);
          tma_load_5d(sa + a_bytes, &tmA, fb, kc * kBK, w1 + tw - a.pw, h1 + th - a.ph, d1 + td - a.pd, g1);
          tma_load_3d(sa + 2 * a_bytes, &tmB, fb, kc * kBK, n0, tap);
          if (++s == a.stages) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_16(kBM, 256, true);
      int s = 0;
      uint32_t ph = 0, item = 0;
      for (int it = blockIdx.x; it < total_items; it += gridDim.x, ++item) {
        mbar_wait(tempty_bar, (item & 1u) ^ 1u, a.err, 2);        // the previous item's epilogue has read TMEM
        tc_fence_after();
        for (int ki = 0; ki < kiters; ++ki) {
          mbar_wait(full_bar(s), ph, a.err, 3);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
          const uint64_t da0 = umma_desc_sw128(sa), da1 = umma_desc_sw128(sa + a_bytes);
          const uint64_t db = umma_desc_sw128(sa + 2 * a_bytes);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            const uint64_t koff = (uint64_t)(k * 2);
            const uint32_t acc = (ki != 0 || k != 0) ? 1u : 0u;
            umma_f16(tmem_base, da0 + koff, db + koff, idesc, acc);
            umma_f16(tmem_base + 256u, da1 + koff, db + koff, idesc, acc);
          }
          umma_commit(empty_bar(s));
          if (++s == a.stages) { s = 0; ph ^= 1u; }
        }
        umma_commit(tfull_bar);
      }
    }
  } else {
    const int q = warp & 3;
    const int colhalf = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    float acc_scale = a.acc_scale;
    if (a.acc_scale_dev0) acc_scale *= *a.acc_scale_dev0;
    if (a.acc_scale_dev1) acc_scale *= *a.acc_scale_dev1;
    uint32_t item = 0;
    for (int it = blockIdx.x; it < total_items; it += gridDim.x, ++item) {
      mbar_wait(tfull_bar, item & 1u, a.err, 4);
      tc_fence_after();
#pragma unroll 1
      for (int which = 0; which < 2; ++which) {
        int n0, w0, h0, d0, g0;
        decode(it, which, n0, w0, h0, d0, g0);
        int rr = row;
        const int ww = w0 + rr % a.bw; rr /= a.bw;
        const int hh = h0 + rr % a.bh; rr /= a.bh;
        const int dd = d0 + rr % a.bd; rr /= a.bd;
        const int gg = g0 + rr;
        const bool row_ok = (ww < a.W) && (hh < a.H) && (dd < a.D) && (gg < a.G);
        const long long m = (((long long)gg * a.D + dd) * a.H + hh) * a.W + ww;
        const long long m_res = a.res_d > 0
                                    ? (((long long)gg * a.res_d + (dd >> 1)) * a.res_h + (hh >> 1)) * a.res_w + (ww >> 1)
                                    : m;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)which * 256u + (uint32_t)(colhalf * 128);
#pragma unroll 1
        for (int b = 0; b < 4; ++b) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(taddr + (uint32_t)(b * 32), v);       // warp-collective: every lane takes part
          tmem_ld_wait();
          const int n = n0 + colhalf * 128 + b * 32;
          if (row_ok && n < a.Cout) {
            float f[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
            epilogue_store32(a, f, m, n, 0, acc_scale, m_res);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// out[m][n] = bias[n] + sum over the K slices, slices added in index order (deterministic).
__global__ void splitk_reduce_kernel(const float* __restrict__ ws, int splits, long long m_total, int cout, long long ld,
                                     const float* __restrict__ bias, float* __restrict__ out) {
  const long long total = m_total * ld;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int n = (int)(i % ld);
    if (n >= cout) continue;
    float acc = bias ? bias[n] : 0.f;
    for (int s = 0; s < splits; ++s) acc += ws[(long long)s * total + i];
    out[i] = acc;
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;

static int ensure_encode() {
  if (g_encode) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || fn == nullptr) {
    set_error("cuTensorMapEncodeTiled is not available from the CUDA driver (%d)", (int)e);
    return DRB_ECUDA;
  }
  g_encode = (PFN_cuTensorMapEncodeTiled_v12000)fn;
  return 0;
}

// 16-bit float tensor map, 128-byte swizzle, zero OOB fill.  dims/box are innermost-first.
int igemm_make_map(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes /* rank-1 */, const uint32_t* box, bool is_bf16) {
  int rc = ensure_encode();
  if (rc) return rc;
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i) gstr[i - 1] = strides_bytes[i - 1];
  }
  CUresult r = g_encode(map, is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, (void*)base, gdim,
                        gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu %llu %llu ...)",
              (int)r, rank, (unsigned long long)dims[0], (unsigned long long)dims[1],
              (unsigned long long)(rank > 2 ? dims[2] : 0));
    return DRB_ECUDA;
  }
  return 0;
}

// Per-device state (a process may drive several GPUs: NeRFRegTr keys its engines by device index).
static constexpr int kMaxDevices = 64;
struct DeviceState {
  int num_sms = 0;
  int* err_flag = nullptr;        // device view of the flag
  int* err_flag_host = nullptr;   // the flag lives in mapped pinned host memory: still readable after a device trap
  bool attr_set = false;
};
static DeviceState g_dev[kMaxDevices];

static DeviceState& device_state() {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= kMaxDevices) dev = 0;
  return g_dev[dev];
}

int igemm_num_sms() {
  DeviceState& st = device_state();
  if (!st.num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&st.num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (st.num_sms <= 0) st.num_sms = 148;
  }
  return st.num_sms;
}

void igemm_clear_err_flag() {
  DeviceState& st = device_state();
  if (st.err_flag_host) memset(st.err_flag_host, 0, 16 * sizeof(int));
}

// Current value of the flag without waiting for anything: it lives in mapped pinned host memory, so whatever a
// kernel of an already synchronised stream wrote is visible.  (drb_igemm_error_flag waits for the whole device.)
int igemm_peek_err_flag() {
  int* flag = device_state().err_flag_host;
  return flag ? *(volatile int*)flag : 0;
}

int* igemm_err_flag() {
  DeviceState& st = device_state();
  if (!st.err_flag) {
    void* h = nullptr;
    if (cudaHostAlloc(&h, 16 * sizeof(int), cudaHostAllocMapped) != cudaSuccess) return nullptr;
    memset(h, 0, 16 * sizeof(int));       // [0] the flag, [1..15] per-site detail of the pipeline watchdogs
    void* d = nullptr;
    if (cudaHostGetDevicePointer(&d, h, 0) != cudaSuccess) { cudaFreeHost(h); return nullptr; }
    st.err_flag_host = (int*)h;
    st.err_flag = (int*)d;
  }
  return st.err_flag;
}

// Chooses the 128-row spatial box (bg, bd, bh, bw) for an output volume.
void igemm_choose_box(int G, int D, int H, int W, int& bg, int& bd, int& bh, int& bw) {
  int rem = kBM;
  auto take = [&](int extent) {
    int b = 1;
    while (b * 2 <= rem && b < extent) b *= 2;
    rem /= b;
    return b;
  };
  // prefer a compact box: cap w and h at 8 first so that 3-D halos overlap in L2
  bw = 1; while (bw * 2 <= 8 && bw < W) bw *= 2;
  if (H == 1 && D == 1) { bw = 1; while (bw * 2 <= kBM && bw < W) bw *= 2; }
  rem /= bw;
  bh = 1; while (bh * 2 <= (rem < 8 ? rem : 8) && bh < H) bh *= 2;
  rem /= bh;
  bd = take(D);
  bg = take(G);
  // whatever is left (tiny volumes) pads the innermost dimension; TMA zero-fills out of bounds rows
  bw *= rem;
}

extern "C" int drb_conv3d_igemm(const drb_conv3d_desc* d, cudaStream_t stream) {
  DRB_REQUIRE(d != nullptr, "drb_conv3d_igemm: null descriptor");
  DRB_REQUIRE(d->planes == 1 || d->planes == 2, "drb_conv3d_igemm: planes must be 1 or 2");
  DRB_REQUIRE(d->cin > 0 && d->cin % kBK == 0, "drb_conv3d_igemm: Cin=%d must be a multiple of 64",
              d->cin);
  DRB_REQUIRE(d->cout > 0, "drb_conv3d_igemm: Cout must be positive");
  DRB_REQUIRE(d->x_hi && d->w_hi, "drb_conv3d_igemm: null operand plane");
  DRB_REQUIRE(d->planes == 1 || (d->x_lo && d->w_lo), "drb_conv3d_igemm: planes==2 needs lo planes");
  DRB_REQUIRE(d->kd >= 1 && d->kh >= 1 && d->kw >= 1 && (d->kd & 1) && (d->kh & 1) && (d->kw & 1),
              "drb_conv3d_igemm: kernel extents must be odd");
  DRB_REQUIRE(d->out || d->out_hi, "drb_conv3d_igemm: no output requested");
  DRB_REQUIRE(!(d->out_lo && !d->out_hi), "drb_conv3d_igemm: out_lo without out_hi");
  const long long ld = d->ld_out > 0 ? d->ld_out : d->cout;

  IgemmArgs a;
  memset(&a, 0, sizeof(a));
  a.G = d->g; a.D = d->d; a.H = d->h; a.W = d->w;
  a.Cin = d->cin; a.Cout = d->cout;
  a.kd = d->kd; a.kh = d->kh; a.kw = d->kw;
  a.pd = d->kd / 2; a.ph = d->kh / 2; a.pw = d->kw / 2;
  igemm_choose_box(a.G, a.D, a.H, a.W, a.bg, a.bd, a.bh, a.bw);
  a.chunk = d->planes == 2 ? 2 : 4;
  const int nsm = igemm_num_sms();
  const long long tiles_m_h = (long long)cdiv(a.W, a.bw) * cdiv(a.H, a.bh) * cdiv(a.D, a.bd) * cdiv(a.G, a.bg);
  // N tile: the widest of 256 / 128 / 64 that still yields at least one tile per SM (small problems are
  // latency bound: more, narrower tiles beat fewer wide ones); never wider than Cout rounded to 64.
  // A narrow tile re-reads the activation box once per N tile (128 B/clk per SM at BN = 64 against an L2 -> SM
  // budget of ~43): when the reduction is long and the epilogue plain, a WIDE tile with the reduction split over the
  // idle SMs (deterministic slices, see below) beats more, narrower tiles (measured: 8192 x 256 x 6912 157 -> ~40 us).
  const int kiters_pre = d->kd * d->kh * d->kw * (d->cin / kBK);
  const bool plain_pre = d->out && !d->out_hi && !d->residual && !d->relu && (d->out_scale == 0.f || d->out_scale == 1.f) &&
                         !d->tile_list && !d->bn_accum;
  static int wide_split = -1;      // DRB_IGEMM_WIDE_SPLIT=0: the round-1 rule (narrow tiles first)
  if (wide_split < 0) { const char* env = getenv("DRB_IGEMM_WIDE_SPLIT"); wide_split = env ? atoi(env) : 1; }
  a.BN = 64;
  for (int bn = 256; bn >= 64; bn >>= 1) {
    if (bn > ((d->cout + 63) / 64) * 64 && bn != 64) continue;
    const long long t = tiles_m_h * cdiv(d->cout, bn);
    if (t >= nsm || bn == 64) { a.BN = bn; break; }
    if (wide_split && plain_pre && d->splitk_ws && bn > 64 && t * 2 <= nsm && kiters_pre >= 8 * a.chunk &&
        2 * (size_t)a.G * a.D * a.H * a.W * (size_t)ld * sizeof(float) <= d->splitk_ws_bytes) {
      a.BN = bn;                   // few wide tiles: the split-K rule below fills the machine
      break;
    }
  }
  if (a.BN > ((d->cout + 63) / 64) * 64) a.BN = ((d->cout + 63) / 64) * 64;
  // split-K: few tiles but a long reduction (deep backbone layers: 1-8 tiles, K up to 13824).  Only for
  // the plain "fp32 out (+ bias)" epilogue; partial sums are added atomically into a zeroed output.
  const int kiters_h = d->kd * d->kh * d->kw * (d->cin / kBK);
  const long long tiles_h = tiles_m_h * cdiv(d->cout, a.BN);
  a.splits = 1;
  a.kper = kiters_h;
  const bool plain = d->out && !d->out_hi && !d->residual && !d->relu && (d->out_scale == 0.f || d->out_scale == 1.f);
  a.m_total = (long long)a.G * a.D * a.H * a.W;
  if (plain && d->splitk_ws && tiles_h * 2 <= nsm && kiters_h >= 4 * a.chunk) {
    int want = (int)(nsm / tiles_h);
    int maxs = kiters_h / (2 * a.chunk);
    int sp = want < maxs ? want : maxs;
    if (sp > 1) {
      int kper = (kiters_h + sp - 1) / sp;
      kper = ((kper + a.chunk - 1) / a.chunk) * a.chunk;
      a.kper = kper;
      a.splits = (kiters_h + kper - 1) / kper;
      // the slices meet in the caller's workspace; too small a workspace simply means no split
      if ((size_t)a.splits * (size_t)a.m_total * (size_t)ld * sizeof(float) > d->splitk_ws_bytes) {
        a.splits = 1;
        a.kper = kiters_h;
      }
    }
  }
  a.splitk_ws = (float*)d->splitk_ws;
  a.tile_list = d->tile_list;
  a.tile_count = d->tile_count;
  // BatchNorm statistics of the output: fused into the epilogue when a tile never straddles two grids
  // and the reduction is not split; otherwise a separate pass over the finished output.
  a.bn_accum = nullptr;
  bool bn_separate = false;
  if (d->bn_accum) {
    DRB_REQUIRE(d->out && !d->bias && !d->residual && !d->relu && !d->tile_list,
                "drb_conv3d_igemm: bn_accum needs the plain fp32-output epilogue");
    DRB_CUDA_OK(cudaMemsetAsync(d->bn_accum, 0, sizeof(double) * 2 * (size_t)a.G * a.Cout, stream));
    if (a.splits == 1 && a.bg == 1) a.bn_accum = d->bn_accum; else bn_separate = true;
  }
  DRB_REQUIRE((d->tile_list == nullptr) == (d->tile_count == nullptr), "drb_conv3d_igemm: tile_list / tile_count pair");
  if (a.tile_list) { a.splits = 1; a.kper = kiters_h; }
  {
    static int no_split = -1;
    if (no_split < 0) { const char* env = getenv("DRB_IGEMM_NO_SPLITK"); no_split = env ? atoi(env) : 0; }
    if (no_split) { a.splits = 1; a.kper = kiters_h; }
  }
  {
    // cluster size along M: multicast pays off when there are enough tiles to keep every CTA busy
    static int forced = -1;
    if (forced < 0) {
      const char* env = getenv("DRB_IGEMM_CLUSTER");
      forced = env ? atoi(env) : 0;
    }
    const long long tm = (long long)cdiv(a.W, a.bw) * cdiv(a.H, a.bh) * cdiv(a.D, a.bd) * cdiv(a.G, a.bg);
    const long long tn = cdiv(a.Cout, a.BN);
    // Measured on B200 (round 1): the large FPN convolutions already run at ~90 % of the bf16 MMA rate
    // with cs = 1 and multicast brings nothing (cs = 2: +-0 %, cs = 4: -30 % from stranded SMs), so the
    // cluster path stays opt-in (DRB_IGEMM_CLUSTER=2|4); it is covered by the parity tests.
    (void)tn;
    a.cs = 1;
    if (forced == 1 || forced == 2 || forced == 4) a.cs = forced;
    if (tm < a.cs || d->tile_list) a.cs = 1;
  }
  a.planes = d->planes;
  a.relu = d->relu;
  a.out_scale = d->out_scale == 0.f ? 1.f : d->out_scale;
  a.acc_scale = d->acc_scale == 0.f ? 1.f : d->acc_scale;
  a.acc_scale_dev0 = d->acc_scale_dev[0];
  a.acc_scale_dev1 = d->acc_scale_dev[1];
  a.res_d = d->res_d; a.res_h = d->res_h; a.res_w = d->res_w;
  DRB_REQUIRE((a.res_d == 0 && a.res_h == 0 && a.res_w == 0) ||
                  (d->residual && a.res_d * 2 >= a.D && a.res_h * 2 >= a.H && a.res_w * 2 >= a.W),
              "drb_conv3d_igemm: the coarse residual volume (%d, %d, %d) does not cover the output (%d, %d, %d) at x2",
              a.res_d, a.res_h, a.res_w, a.D, a.H, a.W);
  a.bias = d->bias; a.residual = d->residual; a.out = d->out; a.out_hi = (plane_t*)d->out_hi; a.out_lo = (plane_t*)d->out_lo;
  a.ld = ld;
  // vector stores in the epilogue need 16-byte aligned rows
  DRB_REQUIRE(ld % 8 == 0, "drb_conv3d_igemm: row pitch %lld must be a multiple of 8 elements", ld);

  const size_t stage_bytes = (size_t)a.planes * ((size_t)kBM * kBK * 2 + (size_t)a.BN * kBK * 2);
  const size_t budget = 200 * 1024;
  int stages = (int)(budget / stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  DRB_REQUIRE(stages >= 2, "drb_conv3d_igemm: tile does not fit shared memory");
  a.stages = stages;
  const size_t smem = 1024 + stages * stage_bytes + (2 * kMaxStages + 4) * 8 + 16;

  a.err = igemm_err_flag();
  DRB_REQUIRE(a.err != nullptr, "drb_conv3d_igemm: could not allocate the device error flag");

  CUtensorMap mA[2], mB[2];
  memset(mA, 0, sizeof(mA));
  memset(mB, 0, sizeof(mB));
  const uint64_t adims[5] = {(uint64_t)a.Cin, (uint64_t)a.W, (uint64_t)a.H, (uint64_t)a.D, (uint64_t)a.G};
  const uint64_t astr[4] = {(uint64_t)a.Cin * 2, (uint64_t)a.W * a.Cin * 2,
                            (uint64_t)a.H * a.W * a.Cin * 2, (uint64_t)a.D * a.H * a.W * a.Cin * 2};
  const uint32_t abox[5] = {(uint32_t)kBK, (uint32_t)a.bw, (uint32_t)a.bh, (uint32_t)a.bd, (uint32_t)a.bg};
  const int taps = a.kd * a.kh * a.kw;
  const uint64_t bdims[3] = {(uint64_t)a.Cin, (uint64_t)a.Cout, (uint64_t)taps};
  const uint64_t bstr[2] = {(uint64_t)a.Cin * 2, (uint64_t)a.Cout * a.Cin * 2};
  const uint32_t bbox[3] = {(uint32_t)kBK, (uint32_t)(a.BN / a.cs), 1u};
  int rc;
  if ((rc = igemm_make_map(&mA[0], d->x_hi, 5, adims, astr, abox, a.planes == 1))) return rc;
  if ((rc = igemm_make_map(&mB[0], d->w_hi, 3, bdims, bstr, bbox, a.planes == 1))) return rc;
  if (a.planes == 2) {
    if ((rc = igemm_make_map(&mA[1], d->x_lo, 5, adims, astr, abox, false))) return rc;
    if ((rc = igemm_make_map(&mB[1], d->w_lo, 3, bdims, bstr, bbox, false))) return rc;
  } else {
    mA[1] = mA[0];
    mB[1] = mB[0];
  }

  bool& attr_set = device_state().attr_set;
  if (!attr_set) {
    DRB_CUDA_OK(cudaFuncSetAttribute(igemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     227 * 1024));
    DRB_CUDA_OK(cudaFuncSetAttribute(igemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     227 * 1024));
    attr_set = true;
  }
  const int tiles_m = cdiv(a.W, a.bw) * cdiv(a.H, a.bh) * cdiv(a.D, a.bd) * cdiv(a.G, a.bg);
  {
    // bf16, wide weight tile, enough M tiles to keep every SM busy with pairs: the two-tile kernel
    static int two = -1;           // DRB_IGEMM_TWO_TILES=0: always the one-tile kernel
    if (two < 0) { const char* env = getenv("DRB_IGEMM_TWO_TILES"); two = env ? atoi(env) : 1; }
    // chosen when the reduction is long enough for the MMAs to dominate an item (the two-tile kernel does not
    // overlap an item's epilogue with the next item's MMAs) and when its waves of (tile pair, N tile) items finish
    // earlier than the one-tile kernel's waves at the measured rates (86 % vs ~50 % of the MMA peak per busy SM)
    const long long nsm2 = igemm_num_sms();
    const long long items1 = (long long)tiles_m * cdiv(a.Cout, 256), items2 = (long long)cdiv(tiles_m, 2) * cdiv(a.Cout, 256);
    const double time1 = (double)((items1 + nsm2 - 1) / nsm2) * 1.0 / 0.5;
    const double time2 = (double)((items2 + nsm2 - 1) / nsm2) * 2.0 / 0.86;
    if (two && a.planes == 1 && a.BN == 256 && a.cs == 1 && a.splits == 1 && !a.bn_accum && a.Cout % 8 == 0 &&
        kiters_h >= 16 && items2 >= nsm2 && time2 < time1) {
      static bool attr2[kMaxDevices] = {};
      int dev = 0;
      cudaGetDevice(&dev);
      if (dev >= 0 && dev < kMaxDevices && !attr2[dev]) {
        DRB_CUDA_OK(cudaFuncSetAttribute(igemm2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr2[dev] = true;
      }
      IgemmArgs a2 = a;
      a2.stages = 3;
      const size_t smem2 = 1024 + 3 * (size_t)(2 * kBM * kBK * 2 + 256 * kBK * 2) + (2 * kMaxStages + 4) * 8 + 16;
      // the weight map's box must be the full 256 rows (cs == 1: it is); A map: the forward's 128-row box
      const long long items = (long long)cdiv(tiles_m, 2) * cdiv(a.Cout, 256);
      const int grid2 = (int)(items < igemm_num_sms() ? items : igemm_num_sms());
      igemm2_kernel<<<grid2, kThreads, smem2, stream>>>(mA[0], mB[0], a2);
      DRB_LAUNCH_OK();
      return 0;
    }
  }
  const int tiles = cdiv(tiles_m, a.cs) * cdiv(a.Cout, a.BN) * a.splits;     // work items (one per cluster)
  const int max_clusters = igemm_num_sms() / a.cs;
  const int clusters = tiles < max_clusters ? tiles : max_clusters;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(clusters * a.cs));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)a.cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // two instantiations: the BatchNorm-statistics epilogue costs registers, the big FPN layers do not pay
  if (a.bn_accum)
    DRB_CUDA_OK(cudaLaunchKernelEx(&cfg, igemm_kernel<true>, mA[0], mA[1], mB[0], mB[1], a));
  else
    DRB_CUDA_OK(cudaLaunchKernelEx(&cfg, igemm_kernel<false>, mA[0], mA[1], mB[0], mB[1], a));
  if (a.splits > 1) {
    const long long total = a.m_total * a.ld;
    int rgrid = (int)((total + 255) / 256);
    if (rgrid > 148 * 8) rgrid = 148 * 8;
    splitk_reduce_kernel<<<rgrid, 256, 0, stream>>>(a.splitk_ws, a.splits, a.m_total, a.Cout, a.ld, a.bias, a.out);
    DRB_LAUNCH_OK();
  }
  if (bn_separate)
    return drb_bn_stats(a.out, a.G, (long long)a.D * a.H * a.W, a.Cout, d->bn_accum, stream);
  return 0;
}

extern "C" int drb_conv3d_tile_shape(int g, int d, int h, int w, int box[4], int tiles[4]) {
  DRB_REQUIRE(box && tiles && g > 0 && d > 0 && h > 0 && w > 0, "drb_conv3d_tile_shape: bad arguments");
  int bg, bd, bh, bw;
  igemm_choose_box(g, d, h, w, bg, bd, bh, bw);
  box[0] = bg; box[1] = bd; box[2] = bh; box[3] = bw;
  tiles[0] = cdiv(g, bg); tiles[1] = cdiv(d, bd); tiles[2] = cdiv(h, bh); tiles[3] = cdiv(w, bw);
  return 0;
}

extern "C" int drb_error_flag_detail(int* host16) {
  if (!host16) return DRB_EINVAL;
  memset(host16, 0, 16 * sizeof(int));
  int* flag = device_state().err_flag_host;
  if (!flag) return 0;
  (void)cudaDeviceSynchronize();
  for (int i = 0; i < 16; ++i) host16[i] = ((volatile int*)flag)[i];
  return 0;
}

extern "C" int drb_error_flag_peek(int* host_value) {
  if (!host_value) return DRB_EINVAL;
  *host_value = igemm_peek_err_flag();
  return 0;
}

extern "C" int drb_error_flag_clear(void) {
  igemm_clear_err_flag();
  return 0;
}

extern "C" int drb_igemm_error_flag(int* host_value) {
  if (!host_value) return DRB_EINVAL;
  *host_value = 0;
  int* flag = device_state().err_flag_host;
  if (!flag) return 0;
  // wait for the work in flight; after a device-side trap this returns an error, but the flag (pinned host
  // memory, written with a system-scope fence before the trap) still tells which watchdog fired
  (void)cudaDeviceSynchronize();
  *host_value = *(volatile int*)flag;
  return 0;
}

}  // namespace drb
