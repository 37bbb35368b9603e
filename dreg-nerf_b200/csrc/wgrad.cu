// Weight gradient of every Conv3d / Linear on the registration path as ONE tcgen05 kernel:
//   dW[tap][co][ci] = sum over voxels m of dY[m][co] * X[m + off(tap)][ci]
// (autograd's conv3d weight gradient for conerf/model/resnet3d.py:81-86,120, feature_pyramid_net.py:24,33
// and the Linear layers of conerf/register/transformer.py:128-138, nerf_regtr.py:268-270).
//
// The reduction runs over voxels, which is the slow axis of the channels-last activations, so both
// operands are fed to the tensor core MN-major: the TMA boxes are exactly the forward's
// [spatial rows][64 channels] boxes (SWIZZLE_128B, zero fill out of bounds == the convolution halo), and the
// shared-memory descriptors declare "64-element MN groups LBO apart, 8-row K groups SBO = 1024 B apart"
// with the a_major / b_major bits of the instruction descriptor set.  No transposed copies of dY or X exist.
//
//   warp 0   : TMA producer - per K step one 64-voxel box of dY (2 x 64 channels) and the tap-shifted box of X
//   warp 1   : TMEM allocator + tcgen05.mma issuer (128 x BN x 16, 4 per box and operand product)
//   warps 2-9: drain each finished chunk from TMEM into fp32 registers (the tensor core accumulates with
//              truncation, see igemm.cu), then red.add the tile into the fp32 gradient in torch's
//              [co][ci][tap] layout
// Work item = (128-wide co tile, BN-wide tile of the flattened (tap, ci) axis, K split): the dY box of a K step
// is shared by every tap, so a narrow layer (Cin = 64 / 128) puts 4 / 2 taps side by side in one 256-wide N tile
// (one dY box + four tap-shifted X boxes per step instead of four steps with a dY box each).  With a tile list
// (output-sparse level-1 FPN convolutions) only the listed 128-voxel tiles are reduced over.
//
// Output: the accumulator row of a thread is 32 consecutive columns of the GEMM's [co][tap * Cin + ci] result;
// torch's layout is [co][ci][tap], i.e. every element of a 3^3 convolution would be a lone 4-byte atomic in its
// own sector (measured: ~65 G atomics/s, 100 us for the 7 M elements of a 512 x 512 x 27 weight whatever the
// voxel count).  With a staging buffer (drb_wgrad_desc.stage) the tile is added with 16-byte vector reductions
// into the GEMM layout and `wgrad_unstage_kernel` transposes it into torch's layout through shared memory;
// 1 x 1 x 1 layers are already in that layout and take the vector path directly.
#include "common.cuh"

#include <stdlib.h>

namespace drb {

static constexpr int kWM = 128;        // co tile = TMEM lanes
static constexpr int kWK = 64;         // voxels per K step (one box)
static constexpr int kWMaxStages = 8;
static constexpr int kWAccWarps = 8;
static constexpr int kWThreads = 64 + kWAccWarps * 32;

struct WgradArgs {
  int G, D, H, W;
  int Cout, Cin;
  int kd, kh, kw, pd, ph, pw;
  int fbg, fbd, fbh, fbw;   // the forward's 128-voxel box (tile numbering of tile_list)
  int bg, bd, bh, bw;       // 64-voxel half box
  int sdim;                 // axis halved to form the half box: 0 g, 1 d, 2 h, 3 w
  int BN, planes, stages, chunk, splits;
  int desc_swap;            // debug: swap LBO / SBO
  const int* tile_list;
  const int* tile_count;
  float scale;
  const float* scale_dev;
  float* out;
  int c_real, taps_real;
  int kflat;                // taps * Cin: extent of the flattened (tap, ci) axis
  int vec;                  // out is [Cout][kflat] (staging buffer or a 1x1x1 layer): 16-byte vector reductions
  int* err;
};

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// MN-major, SWIZZLE_128B shared-memory matrix descriptor: rows (K) of 128 B = 64 MN elements, 8-row groups
// `sbo` bytes apart, the next 64 MN elements `lbo` bytes away.
__device__ __forceinline__ uint64_t umma_desc_sw128_mn(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void __launch_bounds__(kWThreads, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
             const __grid_constant__ CUtensorMap tmB0, const __grid_constant__ CUtensorMap tmB1,
             const WgradArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const uint32_t sub_bytes = kWK * 64 * 2;                 // one [64 voxels][64 channels] box: 8 KB
  const uint32_t a_bytes = 2 * sub_bytes;                  // 128 co
  const uint32_t b_bytes = (uint32_t)(a.BN / 64) * sub_bytes;
  const uint32_t stage_bytes = (uint32_t)a.planes * (a_bytes + b_bytes);

  uint64_t* bars = (uint64_t*)(smem + (size_t)a.stages * stage_bytes);
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (kWMaxStages + s); };
  auto tfull_bar = [&](int s) { return bar0 + 8u * (2 * kWMaxStages + s); };
  auto tempty_bar = [&](int s) { return bar0 + 8u * (2 * kWMaxStages + 2 + s); };
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * kWMaxStages + 4);

  const int ftw = (a.W + a.fbw - 1) / a.fbw, fth = (a.H + a.fbh - 1) / a.fbh;
  const int ftd = (a.D + a.fbd - 1) / a.fbd, ftg = (a.G + a.fbg - 1) / a.fbg;
  const int ftiles = ftw * fth * ftd * ftg;
  const int ntiles = a.tile_list ? *a.tile_count : ftiles;
  const int nboxes = 2 * ntiles;
  const int per = (nboxes + a.splits - 1) / a.splits;
  const int tiles_mo = (a.Cout + kWM - 1) / kWM;
  const int tiles_n = (a.kflat + a.BN - 1) / a.BN;
  const int total_items = tiles_mo * tiles_n * a.splits;
  const uint32_t tmem_cols = (2 * a.BN <= 128) ? 128 : (2 * a.BN <= 256) ? 256 : 512;

  if (threadIdx.x == 0) {
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), kWAccWarps * 32);
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
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto decode_item = [&](int it, int& mt, int& nt, int& j0, int& j1) {
    const int sp = it % a.splits;
    const int r = it / a.splits;
    nt = r % tiles_n;
    mt = r / tiles_n;
    j0 = sp * per;
    j1 = min(nboxes, j0 + per);
  };
  auto decode_box = [&](int j, int& w0, int& h0, int& d0, int& g0) {
    int m = a.tile_list ? a.tile_list[j >> 1] : (j >> 1);
    const int u = j & 1;
    const int tw = m % ftw; m /= ftw;
    const int th = m % fth; m /= fth;
    const int td = m % ftd; m /= ftd;
    w0 = tw * a.fbw; h0 = th * a.fbh; d0 = td * a.fbd; g0 = m * a.fbg;
    if (u) {
      if (a.sdim == 0) g0 += a.bg;
      else if (a.sdim == 1) d0 += a.bd;
      else if (a.sdim == 2) h0 += a.bh;
      else w0 += a.bw;
    }
  };

  if (warp == 0) {
    // ------------------------------- TMA producer -------------------------------------------
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int it = blockIdx.x; it < total_items; it += gridDim.x) {
        int mt, nt, j0, j1;
        decode_item(it, mt, nt, j0, j1);
        // the 64-wide sub-boxes of this N tile: (first channel, tap offset); columns past the end of the
        // flattened axis re-load the last valid sub-box (the epilogue drops them)
        int vc[4], vw[4], vh[4], vd[4];
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          int kl = nt * a.BN + 64 * v;
          if (kl >= a.kflat) kl = a.kflat - 64;
          const int tap = kl / a.Cin;
          vc[v] = kl - tap * a.Cin;
          vw[v] = tap % a.kw - a.pw;
          vh[v] = (tap / a.kw) % a.kh - a.ph;
          vd[v] = tap / (a.kw * a.kh) - a.pd;
        }
        for (int j = j0; j < j1; ++j) {
          int w0, h0, d0, g0;
          decode_box(j, w0, h0, d0, g0);
          mbar_wait(empty_bar(s), ph ^ 1u, a.err, 11);
          const uint32_t fb = full_bar(s);
          mbar_expect_tx(fb, stage_bytes);
          const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
          const uint32_t sb = sa + (uint32_t)a.planes * a_bytes;
          for (int p = 0; p < a.planes; ++p) {
            const CUtensorMap* mA = p ? &tmA1 : &tmA0;
            const CUtensorMap* mB = p ? &tmB1 : &tmB0;
#pragma unroll
            for (int u = 0; u < 2; ++u)
              tma_load_5d(sa + p * a_bytes + u * sub_bytes, mA, fb, mt * kWM + 64 * u, w0, h0, d0, g0);
#pragma unroll
            for (int v = 0; v < 4; ++v)
              if (v * 64 < a.BN)
                tma_load_5d(sb + p * b_bytes + v * sub_bytes, mB, fb, vc[v], w0 + vw[v], h0 + vh[v], d0 + vd[v], g0);
          }
          if (++s == a.stages) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer ----------------------------------------------
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_16(kWM, a.BN, a.planes == 1) | (1u << 15) | (1u << 16);   // A, B MN-major
      const uint32_t lbo = a.desc_swap ? 1024u : sub_bytes;
      const uint32_t sbo = a.desc_swap ? sub_bytes : 1024u;
      int s = 0;
      uint32_t ph = 0;
      uint32_t cc = 0;
      for (int it = blockIdx.x; it < total_items; it += gridDim.x) {
        int mt, nt, j0, j1;
        decode_item(it, mt, nt, j0, j1);
        for (int k0 = j0; k0 < j1; k0 += a.chunk, ++cc) {
          const uint32_t r = cc & 1u, rph = (cc >> 1) & 1u;
          mbar_wait(tempty_bar(r), rph ^ 1u, a.err, 12);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + r * (uint32_t)a.BN;
          const int kend = min(k0 + a.chunk, j1);
          for (int ki = k0; ki < kend; ++ki) {
            mbar_wait(full_bar(s), ph, a.err, 13);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
            const uint32_t sb = sa + (uint32_t)a.planes * a_bytes;
#pragma unroll
            for (int k = 0; k < kWK / 16; ++k) {
              const uint32_t koff = (uint32_t)k * 16u * 128u;          // 16 voxel rows of 128 B
              const uint64_t da0 = umma_desc_sw128_mn(sa + koff, lbo, sbo);
              const uint64_t db0 = umma_desc_sw128_mn(sb + koff, lbo, sbo);
              umma_f16(tmem_d, da0, db0, idesc, (ki != k0 || k != 0) ? 1u : 0u);
              if (a.planes == 2) {
                const uint64_t da1 = umma_desc_sw128_mn(sa + a_bytes + koff, lbo, sbo);
                const uint64_t db1 = umma_desc_sw128_mn(sb + b_bytes + koff, lbo, sbo);
                umma_f16(tmem_d, da0, db1, idesc, 1u);
                umma_f16(tmem_d, da1, db0, idesc, 1u);
              }
            }
            umma_commit(empty_bar(s));
            if (++s == a.stages) { s = 0; ph ^= 1u; }
          }
          umma_commit(tfull_bar(r));
        }
      }
    }
  } else {
    // ------------------------------- accumulate + epilogue ------------------------------------
    const int q = warp & 3;
    const int colhalf = (warp - 2) >> 2;
    const int half = a.BN >> 1;
    const int row = q * 32 + lane;           // accumulator row == co within the tile
    float sc = a.scale;
    if (a.scale_dev) sc *= *a.scale_dev;
    float acc[128];
    uint32_t cc = 0;
    for (int it = blockIdx.x; it < total_items; it += gridDim.x) {
      int mt, nt, j0, j1;
      decode_item(it, mt, nt, j0, j1);
      if (j0 >= j1) continue;
#pragma unroll
      for (int j = 0; j < 128; ++j) acc[j] = 0.f;
      for (int k0 = j0; k0 < j1; k0 += a.chunk, ++cc) {
        const uint32_t r = cc & 1u, rph = (cc >> 1) & 1u;
        mbar_wait(tfull_bar(r), rph, a.err, 14);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + r * (uint32_t)a.BN + (uint32_t)(colhalf * half);
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          if (b * 32 < half) {
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
      const int co = mt * kWM + row;
      if (co < a.Cout) {
        const int kl0 = nt * a.BN + colhalf * half;
        if (a.vec) {
          float* dst = a.out + (long long)co * a.kflat;
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            if (b * 32 < half) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const int kl = kl0 + b * 32 + j;
                if (kl < a.kflat)
                  red_add_v4(dst + kl, acc[b * 32 + j] * sc, acc[b * 32 + j + 1] * sc, acc[b * 32 + j + 2] * sc,
                             acc[b * 32 + j + 3] * sc);
              }
            }
          }
        } else {
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            if (b * 32 < half) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const int klin = kl0 + b * 32 + j;
                if (klin < a.kflat) {
                  const int tp = klin / a.c_real;
                  const int c = klin - tp * a.c_real;
                  if (tp < a.taps_real)
                    atomicAdd(a.out + ((long long)co * a.c_real + c) * a.taps_real + tp, acc[b * 32 + j] * sc);
                }
              }
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// stage [Cout][kflat] (GEMM layout, k = tap * c_real + c) -> out [Cout][c_real][taps_real] += ; one block per
// (32 channels, co): coalesced 128-byte reads per tap, one contiguous run of 32 * taps_real floats written.
__global__ void __launch_bounds__(256)
wgrad_unstage_kernel(const float* __restrict__ stage, int kflat, int c_real, int taps_real, float* __restrict__ out) {
  extern __shared__ float tile[];                 // [taps_real][33]
  const int co = blockIdx.y;
  const int c0 = blockIdx.x * 32;
  const int cw = min(32, c_real - c0);
  const float* srow = stage + (long long)co * kflat + c0;
  for (int i = threadIdx.x; i < taps_real * 32; i += blockDim.x) {
    const int tp = i >> 5, c = i & 31;
    if (c < cw) tile[tp * 33 + c] = srow[(long long)tp * c_real + c];
  }
  __syncthreads();
  float* orow = out + ((long long)co * c_real + c0) * taps_real;
  for (int i = threadIdx.x; i < cw * taps_real; i += blockDim.x) {
    const int c = i / taps_real;
    const int tp = i - c * taps_real;
    orow[i] += tile[tp * 33 + c];
  }
}

extern "C" int drb_conv3d_wgrad(const drb_wgrad_desc* d, cudaStream_t stream) {
  DRB_REQUIRE(d != nullptr, "drb_conv3d_wgrad: null descriptor");
  DRB_REQUIRE(d->planes == 1 || d->planes == 2, "drb_conv3d_wgrad: planes must be 1 or 2");
  DRB_REQUIRE(d->cin > 0 && d->cin % 64 == 0, "drb_conv3d_wgrad: Cin=%d must be a multiple of 64", d->cin);
  DRB_REQUIRE(d->cout > 0 && d->cout % 8 == 0, "drb_conv3d_wgrad: Cout=%d must be a multiple of 8", d->cout);
  DRB_REQUIRE(d->dy_hi && d->x_hi && d->dw, "drb_conv3d_wgrad: null operand");
  DRB_REQUIRE(d->planes == 1 || (d->dy_lo && d->x_lo), "drb_conv3d_wgrad: planes==2 needs lo planes");
  DRB_REQUIRE(d->kd >= 1 && d->kh >= 1 && d->kw >= 1 && (d->kd & 1) && (d->kh & 1) && (d->kw & 1),
              "drb_conv3d_wgrad: kernel extents must be odd");
  DRB_REQUIRE((d->tile_list == nullptr) == (d->tile_count == nullptr), "drb_conv3d_wgrad: tile_list / tile_count pair");

  WgradArgs a;
  memset(&a, 0, sizeof(a));
  a.G = d->g; a.D = d->d; a.H = d->h; a.W = d->w;
  a.Cout = d->cout; a.Cin = d->cin;
  a.kd = d->kd; a.kh = d->kh; a.kw = d->kw;
  a.pd = d->kd / 2; a.ph = d->kh / 2; a.pw = d->kw / 2;
  igemm_choose_box(a.G, a.D, a.H, a.W, a.fbg, a.fbd, a.fbh, a.fbw);
  a.bg = a.fbg; a.bd = a.fbd; a.bh = a.fbh; a.bw = a.fbw;
  if (a.fbg > 1) { a.sdim = 0; a.bg = a.fbg / 2; }
  else if (a.fbd > 1) { a.sdim = 1; a.bd = a.fbd / 2; }
  else if (a.fbh > 1) { a.sdim = 2; a.bh = a.fbh / 2; }
  else { a.sdim = 3; a.bw = a.fbw / 2; }
  a.planes = d->planes;
  a.chunk = d->planes == 2 ? 2 : 4;
  const int taps = a.kd * a.kh * a.kw;
  a.c_real = d->c_real > 0 ? d->c_real : d->cin;
  a.taps_real = d->taps_real > 0 ? d->taps_real : taps;
  DRB_REQUIRE((long long)a.c_real * a.taps_real <= (long long)taps * d->cin,
              "drb_conv3d_wgrad: c_real * taps_real exceeds the GEMM's K extent");
  a.scale = d->scale == 0.f ? 1.f : d->scale;
  a.scale_dev = d->scale_dev;
  a.out = d->dw;
  DRB_REQUIRE((long long)taps * d->cin < (1LL << 31), "drb_conv3d_wgrad: taps * Cin too large");
  a.kflat = taps * d->cin;
  // Output path: the GEMM layout [Cout][kflat] IS torch's layout for a plain 1x1x1 layer; otherwise go through the
  // caller's staging buffer when it is large enough (else: scattered scalar atomics, correct but slow).
  const bool same_layout = a.taps_real == 1 && a.c_real == d->cin && taps == 1;
  const long long stage_need = (long long)d->cout * a.kflat;
  const size_t unstage_smem = (size_t)a.taps_real * 33 * sizeof(float);
  static int use_stage = -1;
  if (use_stage < 0) { const char* env = getenv("DRB_WGRAD_STAGE"); use_stage = env ? atoi(env) : 1; }
  const bool staged = !same_layout && use_stage && d->stage != nullptr && d->stage_elems >= stage_need &&
                      unstage_smem <= 48 * 1024 && ((uintptr_t)d->stage & 15) == 0;
  if (staged) {
    a.out = d->stage;
    a.vec = 1;
    DRB_CUDA_OK(cudaMemsetAsync(d->stage, 0, (size_t)stage_need * sizeof(float), stream));
  } else if (same_layout && ((uintptr_t)d->dw & 15) == 0) {
    a.vec = 1;
  }
  a.tile_list = d->tile_list;
  a.tile_count = d->tile_count;
  {
    static int swap = -1;
    if (swap < 0) { const char* env = getenv("DRB_WGRAD_DESC_SWAP"); swap = env ? atoi(env) : 0; }
    a.desc_swap = swap;
  }
  const int nsm = igemm_num_sms();
  const long long ftiles = (long long)cdiv(a.W, a.fbw) * cdiv(a.H, a.fbh) * cdiv(a.D, a.fbd) * cdiv(a.G, a.fbg);
  const long long nboxes = 2 * ftiles;
  // Work decomposition: (128 co, BN columns of the flattened (tap, ci) axis) tiles x K splits.  Small problems (a few hundred tokens, deep
  // backbone levels) are latency bound: prefer narrower ci tiles and finer K splits until every SM has an item;
  // large ones (level-1 FPN: thousands of boxes) take the widest tile and ~2 items per SM.
  long long max_splits = nboxes / a.chunk;
  if (max_splits < 1) max_splits = 1;
  long long base_items = 0;
  static int flat_n = -1;       // DRB_WGRAD_FLAT_N=0: N tiles never span taps (the round-2 first version)
  if (flat_n < 0) { const char* env = getenv("DRB_WGRAD_FLAT_N"); flat_n = env ? atoi(env) : 1; }
  const int n_extent = (flat_n || d->cin % 256 == 0) ? a.kflat : d->cin;
  for (int bn = 256; bn >= 64; bn >>= 1) {
    if ((bn > n_extent || (!flat_n && d->cin % bn != 0)) && bn != 64) continue;
    a.BN = bn;
    base_items = (long long)cdiv(a.Cout, kWM) * cdiv(a.kflat, bn);
    if (base_items * max_splits >= nsm) break;
  }
  // K splits from a wave model: the persistent CTAs stride over the items, so a launch lasts
  // ceil(items / SMs) x (boxes per item + the item's fixed cost: pipeline fill, TMEM drain, reductions ~ 6 boxes).
  // "About two items per SM" (the first rule) often meant 2.01-2.2 waves, i.e. three items on a few SMs while the
  // rest idle: 324 items of 171 boxes for 256 x 256 x 27 at 2 x 32^3 where 432 items of 128 boxes finish 24 % sooner.
  static int split_model = -1;   // DRB_WGRAD_SPLIT_MODEL=0: the first rule
  if (split_model < 0) { const char* env = getenv("DRB_WGRAD_SPLIT_MODEL"); split_model = env ? atoi(env) : 1; }
  long long splits = (2LL * nsm + base_items - 1) / base_items;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  if (split_model) {
    const double item_cost = 6.0;
    double best = 1e300;
    const long long s_hi = max_splits < 4096 ? max_splits : 4096;
    for (long long s = 1; s <= s_hi; ++s) {
      const long long waves = (base_items * s + nsm - 1) / nsm;
      const long long per_s = (nboxes + s - 1) / s;
      const double cost = (double)waves * ((double)per_s + item_cost);
      if (cost < best * 0.995) { best = cost; splits = s; }     // ties: the fewer splits (less reduction traffic)
    }
  }
  a.splits = (int)splits;

  const size_t stage_bytes = (size_t)a.planes * ((size_t)kWM * kWK * 2 + (size_t)a.BN * kWK * 2);
  const size_t budget = 200 * 1024;
  int stages = (int)(budget / stage_bytes);
  if (stages > kWMaxStages) stages = kWMaxStages;
  DRB_REQUIRE(stages >= 2, "drb_conv3d_wgrad: tile does not fit shared memory");
  a.stages = stages;
  const size_t smem = 1024 + stages * stage_bytes + (2 * kWMaxStages + 4) * 8 + 16;
  a.err = igemm_err_flag();
  DRB_REQUIRE(a.err != nullptr, "drb_conv3d_wgrad: could not allocate the device error flag");

  CUtensorMap mA[2], mB[2];
  memset(mA, 0, sizeof(mA));
  memset(mB, 0, sizeof(mB));
  const uint64_t adims[5] = {(uint64_t)a.Cout, (uint64_t)a.W, (uint64_t)a.H, (uint64_t)a.D, (uint64_t)a.G};
  const uint64_t astr[4] = {(uint64_t)a.Cout * 2, (uint64_t)a.W * a.Cout * 2, (uint64_t)a.H * a.W * a.Cout * 2,
                            (uint64_t)a.D * a.H * a.W * a.Cout * 2};
  const uint64_t bdims[5] = {(uint64_t)a.Cin, (uint64_t)a.W, (uint64_t)a.H, (uint64_t)a.D, (uint64_t)a.G};
  const uint64_t bstr[4] = {(uint64_t)a.Cin * 2, (uint64_t)a.W * a.Cin * 2, (uint64_t)a.H * a.W * a.Cin * 2,
                            (uint64_t)a.D * a.H * a.W * a.Cin * 2};
  const uint32_t box[5] = {64u, (uint32_t)a.bw, (uint32_t)a.bh, (uint32_t)a.bd, (uint32_t)a.bg};
  int rc;
  if ((rc = igemm_make_map(&mA[0], d->dy_hi, 5, adims, astr, box, a.planes == 1))) return rc;
  if ((rc = igemm_make_map(&mB[0], d->x_hi, 5, bdims, bstr, box, a.planes == 1))) return rc;
  if (a.planes == 2) {
    if ((rc = igemm_make_map(&mA[1], d->dy_lo, 5, adims, astr, box, false))) return rc;
    if ((rc = igemm_make_map(&mB[1], d->x_lo, 5, bdims, bstr, box, false))) return rc;
  } else {
    mA[1] = mA[0];
    mB[1] = mB[0];
  }
  DRB_CUDA_OK(cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  const long long items = base_items * a.splits;
  const int grid = (int)(items < nsm ? items : nsm);
  wgrad_kernel<<<grid, kWThreads, smem, stream>>>(mA[0], mA[1], mB[0], mB[1], a);
  DRB_LAUNCH_OK();
  if (staged) {
    const dim3 ug((unsigned)cdiv(a.c_real, 32), (unsigned)d->cout);
    wgrad_unstage_kernel<<<ug, 256, unstage_smem, stream>>>(d->stage, a.kflat, a.c_real, a.taps_real, d->dw);
    DRB_LAUNCH_OK();
  }
  return 0;
}

}  // namespace drb
