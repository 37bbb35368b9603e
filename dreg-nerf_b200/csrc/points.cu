// R3 (trilinear-at-mask gather), R4 (hierarchical voxel-average down-sampling), R5 (sine
// positional embedding).  HBM-bound: coalesced 128-bit row reads, one warp per output row.
#include "common.cuh"

#include <cub/cub.cuh>

namespace drb {

// ------------------------------------------------------------------------------------------
// R3.  The reference materialises F.interpolate(p1, size=(Z,X,Y), trilinear, align_corners=True)
// (2.15 GB at 128^3) and then indexes K rows (nerf_regtr.py:138-147).  Here only the K masked rows
// are interpolated: one warp per row, 8 corner rows of c floats read with float4 loads.
// Lerp nesting follows ATen's upsample_trilinear3d: d(h(w)).
// ------------------------------------------------------------------------------------------
__global__ void trilinear_gather_kernel(const float* __restrict__ p1, int dc, int hc, int wc, int c,
                                        const float* __restrict__ grid, long long s_ch,
                                        long long s_z, long long s_x, long long s_y, int X, int Y,
                                        int Z, const long long* __restrict__ mask, int k,
                                        float* __restrict__ rows, int ld, int* __restrict__ err) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= k) return;
  const long long idx = mask[warp];
  if (idx < 0 || idx >= (long long)X * Y * Z) {      // stale mask of another resolution: flag, do not touch memory
    if (lane == 0 && err) *(volatile int*)err = 21;
    return;
  }
  const int z = (int)(idx % Z);
  const int y = (int)((idx / Z) % Y);
  const int x = (int)(idx / ((long long)Z * Y));
  // conv volume axes: d = z, h = x, w = y
  const float sd = Z > 1 ? (float)(dc - 1) / (float)(Z - 1) : 0.f;
  const float sh = X > 1 ? (float)(hc - 1) / (float)(X - 1) : 0.f;
  const float sw = Y > 1 ? (float)(wc - 1) / (float)(Y - 1) : 0.f;
  const float fd = sd * (float)z, fh = sh * (float)x, fw = sw * (float)y;
  const int d0 = (int)fd, h0 = (int)fh, w0 = (int)fw;
  const int d1 = d0 + (d0 < dc - 1 ? 1 : 0), h1 = h0 + (h0 < hc - 1 ? 1 : 0),
            w1 = w0 + (w0 < wc - 1 ? 1 : 0);
  const float ld1 = fd - (float)d0, ld0 = 1.f - ld1;
  const float lh1 = fh - (float)h0, lh0 = 1.f - lh1;
  const float lw1 = fw - (float)w0, lw0 = 1.f - lw1;
  auto rowp = [&](int d, int h, int w) { return p1 + (((long long)d * hc + h) * wc + w) * c; };
  const float* r000 = rowp(d0, h0, w0); const float* r001 = rowp(d0, h0, w1);
  const float* r010 = rowp(d0, h1, w0); const float* r011 = rowp(d0, h1, w1);
  const float* r100 = rowp(d1, h0, w0); const float* r101 = rowp(d1, h0, w1);
  const float* r110 = rowp(d1, h1, w0); const float* r111 = rowp(d1, h1, w1);
  float* out = rows + (long long)warp * ld;
  if (lane < 4) {
    float v = 0.f;
    if (lane < 3) v = grid[lane * s_ch + z * s_z + x * s_x + y * s_y];
    out[lane] = v;
  }
  for (int cc = lane * 4; cc < c; cc += 128) {
    const float4 a000 = *(const float4*)(r000 + cc), a001 = *(const float4*)(r001 + cc);
    const float4 a010 = *(const float4*)(r010 + cc), a011 = *(const float4*)(r011 + cc);
    const float4 a100 = *(const float4*)(r100 + cc), a101 = *(const float4*)(r101 + cc);
    const float4 a110 = *(const float4*)(r110 + cc), a111 = *(const float4*)(r111 + cc);
    float4 o;
#define DRB_TRI(f)                                                                          \
  o.f = ld0 * (lh0 * (lw0 * a000.f + lw1 * a001.f) + lh1 * (lw0 * a010.f + lw1 * a011.f)) + \
        ld1 * (lh0 * (lw0 * a100.f + lw1 * a101.f) + lh1 * (lw0 * a110.f + lw1 * a111.f));
    DRB_TRI(x) DRB_TRI(y) DRB_TRI(z) DRB_TRI(w)
#undef DRB_TRI
    *(float4*)(out + 4 + cc) = o;
  }
}

extern "C" int drb_trilinear_gather(const float* p1, int dc, int hc, int wc, int c,
                                    const float* grid, long long s_ch, long long s_z, long long s_x,
                                    long long s_y, int X, int Y, int Z, const long long* mask, int k,
                                    float* rows_out, int ld_rows, cudaStream_t stream) {
  DRB_REQUIRE(p1 && grid && mask && rows_out, "drb_trilinear_gather: null argument");
  DRB_REQUIRE(c % 4 == 0 && ld_rows >= 4 + c && ld_rows % 4 == 0, "drb_trilinear_gather: bad pitch");
  if (k == 0) return 0;
  const int warps_per_block = 8;
  trilinear_gather_kernel<<<cdiv(k, warps_per_block), warps_per_block * 32, 0, stream>>>(
      p1, dc, hc, wc, c, grid, s_ch, s_z, s_x, s_y, X, Y, Z, mask, k, rows_out, ld_rows, igemm_err_flag());
  DRB_LAUNCH_OK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// Output-sparse FPN support.  p1 is only ever read by the gather above, so the two convolutions that
// hold 84 % of the FLOPs (pyramid_transformation_1, upsample_transform_1) need only the 128-voxel output
// tiles that contain a voxel the gather touches (plus a one-voxel halo for the layer before).
// mark: need[d][h][w] = 1 for the 8 trilinear corners of every masked voxel (same index arithmetic as
// trilinear_gather_kernel).  list: tiles whose (optionally 1-dilated) box contains a needed voxel.
// ------------------------------------------------------------------------------------------
__global__ void mark_need_kernel(const long long* __restrict__ mask, int k, int X, int Y, int Z, int dc,
                                 int hc, int wc, uint8_t* __restrict__ need, int* __restrict__ err) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= k) return;
  const long long idx = mask[i];
  if (idx < 0 || idx >= (long long)X * Y * Z) {
    if (err) *(volatile int*)err = 21;
    return;
  }
  const int z = (int)(idx % Z);
  const int y = (int)((idx / Z) % Y);
  const int x = (int)(idx / ((long long)Z * Y));
  const float sd = Z > 1 ? (float)(dc - 1) / (float)(Z - 1) : 0.f;
  const float sh = X > 1 ? (float)(hc - 1) / (float)(X - 1) : 0.f;
  const float sw = Y > 1 ? (float)(wc - 1) / (float)(Y - 1) : 0.f;
  const int d0 = (int)(sd * (float)z), h0 = (int)(sh * (float)x), w0 = (int)(sw * (float)y);
  const int d1 = d0 + (d0 < dc - 1 ? 1 : 0), h1 = h0 + (h0 < hc - 1 ? 1 : 0), w1 = w0 + (w0 < wc - 1 ? 1 : 0);
  const int ds[2] = {d0, d1}, hs[2] = {h0, h1}, ws[2] = {w0, w1};
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
      for (int c = 0; c < 2; ++c) need[((long long)ds[a] * hc + hs[b]) * wc + ws[c]] = 1;
}

__global__ void tile_list_kernel(const uint8_t* __restrict__ need, int g, int d, int h, int w, int bg, int bd,
                                 int bh, int bw, int tg, int td, int th, int tw, int dilate,
                                 int* __restrict__ list, int* __restrict__ count,
                                 unsigned long long* __restrict__ total) {
  const int t = blockIdx.x;                     // one block per tile
  if (t >= tg * td * th * tw) return;
  int r = t;
  const int iw = r % tw; r /= tw;
  const int ih = r % th; r /= th;
  const int id = r % td; r /= td;
  const int ig = r;
  const int ew = bw + 2 * dilate, eh = bh + 2 * dilate, ed = bd + 2 * dilate;
  const int vox = bg * ed * eh * ew;
  int any = 0;
  for (int v = threadIdx.x; v < vox && !any; v += blockDim.x) {
    int q = v;
    const int xw = iw * bw - dilate + q % ew; q /= ew;
    const int yh = ih * bh - dilate + q % eh; q /= eh;
    const int zd = id * bd - dilate + q % ed; q /= ed;
    const int gg = ig * bg + q;
    if (xw >= 0 && xw < w && yh >= 0 && yh < h && zd >= 0 && zd < d && gg < g)
      any |= need[(((long long)gg * d + zd) * h + yh) * w + xw];
  }
  if (__syncthreads_or(any) && threadIdx.x == 0) {
    list[atomicAdd(count, 1)] = t;
    if (total) atomicAdd(total, 1ull);          // running total for the roofline bookkeeping
  }
}

// need: uint8 [g][dc][hc][wc] (zeroed by the call), masks of the g grids, tile lists for the conv
// output (dilate 0) and for its input side (dilate 1).  counts: int[2] device.
extern "C" int drb_fpn_need_tiles(const long long* const* masks_host, const int* ks_host, int g, int X, int Y,
                                  int Z, int dc, int hc, int wc, uint8_t* need, int* list_out, int* list_in,
                                  int* counts, unsigned long long* totals, cudaStream_t stream) {
  DRB_REQUIRE(masks_host && ks_host && need && list_out && list_in && counts && g > 0,
              "drb_fpn_need_tiles: null argument");
  const long long vol = (long long)dc * hc * wc;
  DRB_CUDA_OK(cudaMemsetAsync(need, 0, (size_t)g * vol, stream));
  DRB_CUDA_OK(cudaMemsetAsync(counts, 0, 2 * sizeof(int), stream));
  for (int i = 0; i < g; ++i) {
    if (ks_host[i] == 0) continue;
    mark_need_kernel<<<cdiv(ks_host[i], 256), 256, 0, stream>>>(masks_host[i], ks_host[i], X, Y, Z, dc, hc, wc,
                                                              need + i * vol, igemm_err_flag());
    DRB_LAUNCH_OK();
  }
  int box[4], tiles[4];
  int rc = drb_conv3d_tile_shape(g, dc, hc, wc, box, tiles);
  if (rc) return rc;
  const int nt = tiles[0] * tiles[1] * tiles[2] * tiles[3];
  tile_list_kernel<<<nt, 128, 0, stream>>>(need, g, dc, hc, wc, box[0], box[1], box[2], box[3], tiles[0],
                                           tiles[1], tiles[2], tiles[3], 0, list_out, counts, totals);
  DRB_LAUNCH_OK();
  tile_list_kernel<<<nt, 128, 0, stream>>>(need, g, dc, hc, wc, box[0], box[1], box[2], box[3], tiles[0],
                                           tiles[1], tiles[2], tiles[3], 1, list_in, counts + 1, totals ? totals + 1 : nullptr);
  DRB_LAUNCH_OK();
  return 0;
}

// Extra tile list for the backward pass: tiles within `dilate` voxels of a needed voxel (need as left by
// drb_fpn_need_tiles).  count: device int, zeroed by the call.
extern "C" int drb_fpn_dilated_tiles(const uint8_t* need, int g, int dc, int hc, int wc, int dilate, int* list,
                                     int* count, unsigned long long* total, cudaStream_t stream) {
  DRB_REQUIRE(need && list && count && g > 0 && dilate >= 0, "drb_fpn_dilated_tiles: bad arguments");
  DRB_CUDA_OK(cudaMemsetAsync(count, 0, sizeof(int), stream));
  int box[4], tiles[4];
  int rc = drb_conv3d_tile_shape(g, dc, hc, wc, box, tiles);
  if (rc) return rc;
  const int nt = tiles[0] * tiles[1] * tiles[2] * tiles[3];
  tile_list_kernel<<<nt, 128, 0, stream>>>(need, g, dc, hc, wc, box[0], box[1], box[2], box[3], tiles[0], tiles[1],
                                           tiles[2], tiles[3], dilate, list, count, total);
  DRB_LAUNCH_OK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// R4.  One round: key = (cloud, cx, cy, cz) packed into 64 bits (21-bit biased cell coordinates),
// stable radix sort of (key, row), segment heads by key change, one warp per output cell averaging
// its member rows in ascending input-row order (the oracle's defined order).
// ------------------------------------------------------------------------------------------
static constexpr int kCellBias = 1 << 20;

__global__ void cell_key_kernel(const float* __restrict__ rows, int ld, int n, int n_src, float dl,
                                unsigned long long* __restrict__ keys, int* __restrict__ vals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* r = rows + (long long)i * ld;
  // IEEE division, as torch's CPU `points / sample_dl` (grid_downsample.py:26)
  const long long cx = (long long)floorf(__fdiv_rn(r[0], dl)) + kCellBias;
  const long long cy = (long long)floorf(__fdiv_rn(r[1], dl)) + kCellBias;
  const long long cz = (long long)floorf(__fdiv_rn(r[2], dl)) + kCellBias;
  const unsigned long long cloud = i >= n_src ? 1ull : 0ull;
  keys[i] = (cloud << 63) | ((unsigned long long)(cx & 0x1FFFFF) << 42) |
            ((unsigned long long)(cy & 0x1FFFFF) << 21) | (unsigned long long)(cz & 0x1FFFFF);
  vals[i] = i;
}

__global__ void head_flag_kernel(const unsigned long long* __restrict__ keys, int n,
                                 int* __restrict__ flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  flags[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

// seg_id[i] = inclusive_scan(flags)[i] - 1.  Writes the start offset of every segment and the
// number of output rows per cloud.
__global__ void segment_start_kernel(const unsigned long long* __restrict__ keys,
                                     const int* __restrict__ flags, const int* __restrict__ scan,
                                     int n, int* __restrict__ seg_start, int* __restrict__ counts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (flags[i]) {
    seg_start[scan[i] - 1] = i;
    atomicAdd(&counts[(int)(keys[i] >> 63)], 1);
  }
  if (i == n - 1) seg_start[scan[i]] = n;   // sentinel end
}

__global__ void segment_mean_kernel(const float* __restrict__ rows, int ld,
                                    const int* __restrict__ sorted_rows,
                                    const int* __restrict__ seg_start, int n_seg,
                                    float* __restrict__ out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= n_seg) return;
  const int s0 = seg_start[warp], s1 = seg_start[warp + 1];
  const float cnt = (float)(s1 - s0);
  for (int cc = lane * 4; cc < ld; cc += 128) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = s0; j < s1; ++j) {
      const float4 v = *(const float4*)(rows + (long long)sorted_rows[j] * ld + cc);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    acc.x = __fdiv_rn(acc.x, cnt); acc.y = __fdiv_rn(acc.y, cnt);
    acc.z = __fdiv_rn(acc.z, cnt); acc.w = __fdiv_rn(acc.w, cnt);
    *(float4*)(out + (long long)warp * ld + cc) = acc;
  }
}

struct DsLayout {
  size_t keys_a, keys_b, vals_a, vals_b, flags, scan, seg_start, counts, rows_a, rows_b, cub, total;
  size_t cub_bytes;
};

static DsLayout ds_layout(int n, int ld) {
  DsLayout L;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += (bytes + 255) & ~(size_t)255;
    return o;
  };
  const size_t nn = (size_t)(n > 0 ? n : 1);
  L.keys_a = take(nn * 8); L.keys_b = take(nn * 8);
  L.vals_a = take(nn * 4); L.vals_b = take(nn * 4);
  L.flags = take(nn * 4); L.scan = take(nn * 4);
  L.seg_start = take((nn + 1) * 4);
  L.counts = take(16);
  L.rows_a = take(nn * ld * 4); L.rows_b = take(nn * ld * 4);
  size_t sort_bytes = 0, scan_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const unsigned long long*)nullptr,
                                  (unsigned long long*)nullptr, (const int*)nullptr, (int*)nullptr,
                                  (int)nn);
  cub::DeviceScan::InclusiveSum(nullptr, scan_bytes, (const int*)nullptr, (int*)nullptr, (int)nn);
  L.cub_bytes = sort_bytes > scan_bytes ? sort_bytes : scan_bytes;
  L.cub = take(L.cub_bytes);
  L.total = off;
  return L;
}

extern "C" size_t drb_downsample_workspace_bytes(int n_rows, int ld) {
  return ds_layout(n_rows, ld).total;
}

extern "C" int drb_hierarchical_downsample_tape(const float* rows, int n_src, int n_tgt, int ld,
                                                int num_rounds, double dl0, int max_total,
                                                void* workspace, size_t workspace_bytes, float* rows_out,
                                                int* host_n_src_out, int* host_n_tgt_out, int* tape,
                                                long long tape_capacity, int* host_round_info,
                                                int* host_rounds, cudaStream_t stream) {
  DRB_REQUIRE(rows && rows_out && workspace && host_n_src_out && host_n_tgt_out,
              "drb_hierarchical_downsample: null argument");
  DRB_REQUIRE(!tape || (host_round_info && host_rounds), "drb_hierarchical_downsample_tape: tape needs the host arrays");
  long long tape_used = 0;
  int rounds_done = 0;
  DRB_REQUIRE(ld % 4 == 0 && ld >= 4, "drb_hierarchical_downsample: pitch must be a multiple of 4");
  const int n0 = n_src + n_tgt;
  const DsLayout L = ds_layout(n0, ld);
  DRB_REQUIRE(workspace_bytes >= L.total, "drb_hierarchical_downsample: workspace too small (%zu < %zu)",
              workspace_bytes, L.total);
  uint8_t* ws = (uint8_t*)workspace;
  auto* keys_a = (unsigned long long*)(ws + L.keys_a);
  auto* keys_b = (unsigned long long*)(ws + L.keys_b);
  int* vals_a = (int*)(ws + L.vals_a);
  int* vals_b = (int*)(ws + L.vals_b);
  int* flags = (int*)(ws + L.flags);
  int* scan = (int*)(ws + L.scan);
  int* seg_start = (int*)(ws + L.seg_start);
  int* counts = (int*)(ws + L.counts);
  float* bufs[2] = {(float*)(ws + L.rows_a), (float*)(ws + L.rows_b)};

  const float* cur = rows;
  int cur_src = n_src, cur_tgt = n_tgt;
  int which = 0;
  double dl = dl0;
  bool any_round = false;
  for (int round = 0; round < num_rounds; ++round) {
    const int n = cur_src + cur_tgt;
    if (n == 0) break;
    const int nb = cdiv(n, 256);
    cell_key_kernel<<<nb, 256, 0, stream>>>(cur, ld, n, cur_src, (float)dl, keys_a, vals_a);
    DRB_LAUNCH_OK();
    size_t cub_bytes = L.cub_bytes;
    DRB_CUDA_OK(cub::DeviceRadixSort::SortPairs(ws + L.cub, cub_bytes, keys_a, keys_b, vals_a,
                                                vals_b, n, 0, 64, stream));
    head_flag_kernel<<<nb, 256, 0, stream>>>(keys_b, n, flags);
    DRB_LAUNCH_OK();
    cub_bytes = L.cub_bytes;
    DRB_CUDA_OK(cub::DeviceScan::InclusiveSum(ws + L.cub, cub_bytes, flags, scan, n, stream));
    DRB_CUDA_OK(cudaMemsetAsync(counts, 0, 2 * sizeof(int), stream));
    segment_start_kernel<<<nb, 256, 0, stream>>>(keys_b, flags, scan, n, seg_start, counts);
    DRB_LAUNCH_OK();
    int host_counts[2] = {0, 0};
    DRB_CUDA_OK(cudaMemcpyAsync(host_counts, counts, 2 * sizeof(int), cudaMemcpyDeviceToHost, stream));
    DRB_CUDA_OK(cudaStreamSynchronize(stream));
    const int n_seg = host_counts[0] + host_counts[1];
    float* dst = bufs[which];
    segment_mean_kernel<<<cdiv(n_seg, 8), 256, 0, stream>>>(cur, ld, vals_b, seg_start, n_seg, dst);
    DRB_LAUNCH_OK();
    if (tape) {
      // record this round for the backward pass: the sort permutation and the segment boundaries
      DRB_REQUIRE(tape_used + n + n_seg + 1 <= tape_capacity, "drb_hierarchical_downsample_tape: tape too small");
      DRB_CUDA_OK(cudaMemcpyAsync(tape + tape_used, vals_b, sizeof(int) * (size_t)n, cudaMemcpyDeviceToDevice, stream));
      DRB_CUDA_OK(cudaMemcpyAsync(tape + tape_used + n, seg_start, sizeof(int) * (size_t)(n_seg + 1),
                                  cudaMemcpyDeviceToDevice, stream));
      host_round_info[2 * rounds_done] = n;
      host_round_info[2 * rounds_done + 1] = n_seg;
      tape_used += (long long)n + n_seg + 1;
      ++rounds_done;
    }
    cur = dst;
    which ^= 1;
    cur_src = host_counts[0];
    cur_tgt = host_counts[1];
    dl *= 2.0;
    any_round = true;
    if (cur_src + cur_tgt <= max_total) break;
  }
  (void)any_round;
  DRB_CUDA_OK(cudaMemcpyAsync(rows_out, cur, (size_t)(cur_src + cur_tgt) * ld * sizeof(float),
                              cudaMemcpyDeviceToDevice, stream));
  *host_n_src_out = cur_src;
  *host_n_tgt_out = cur_tgt;
  if (host_rounds) *host_rounds = rounds_done;
  return 0;
}

extern "C" int drb_hierarchical_downsample(const float* rows, int n_src, int n_tgt, int ld,
                                           int num_rounds, double dl0, int max_total,
                                           void* workspace, size_t workspace_bytes, float* rows_out,
                                           int* host_n_src_out, int* host_n_tgt_out,
                                           cudaStream_t stream) {
  return drb_hierarchical_downsample_tape(rows, n_src, n_tgt, ld, num_rounds, dl0, max_total, workspace,
                                          workspace_bytes, rows_out, host_n_src_out, host_n_tgt_out, nullptr, 0,
                                          nullptr, nullptr, stream);
}

// ------------------------------------------------------------------------------------------
// R5: out[i][a*84 + 2j] = sin(p_a / T^(2j/84)), [.. + 2j+1] = cos(p_a / T^(2j/84)), 4 zero pads.
// ------------------------------------------------------------------------------------------
__global__ void pos_embed_kernel(const float* __restrict__ xyz, int ld_xyz, int n, float scale2pi,
                                 float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)n * 256) return;
  const int col = (int)(i & 255);
  const int row = (int)(i >> 8);
  float v = 0.f;
  if (col < 252) {
    const int axis = col / 84;
    const int f = col % 84;
    // temperature ** (2 * (f // 2) / 84), fp32 like torch (position_embedding.py:40-43)
    const float expo = (float)(2 * (f / 2)) / 84.f;
    const float dim_t = powf(1000.f, expo);
    const float arg = __fdiv_rn(xyz[(long long)row * ld_xyz + axis] * scale2pi, dim_t);
    v = (f & 1) ? cosf(arg) : sinf(arg);
  }
  out[i] = v;
}

extern "C" int drb_pos_embed_sine(const float* xyz, int ld_xyz, int n, float scale, float* out,
                                  cudaStream_t stream) {
  DRB_REQUIRE(xyz && out && ld_xyz >= 3, "drb_pos_embed_sine: bad arguments");
  if (n == 0) return 0;
  const float scale2pi = (float)((double)scale * 2.0 * 3.14159265358979323846);
  pos_embed_kernel<<<cdiv((long long)n * 256, 256), 256, 0, stream>>>(xyz, ld_xyz, n, scale2pi, out);
  DRB_LAUNCH_OK();
  return 0;
}

}  // namespace drb
